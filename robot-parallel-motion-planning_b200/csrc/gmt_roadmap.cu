// gmt_roadmap.cu -- roadmap builders for the planner: counter-based state sampler and
// radius neighbour search over a uniform cell grid (sm_100a).
//
// They replace, for the synthetic 2-D world, what the reference does on the CPU in
// CudaAgent.create_samples (carla/cuda_agent.py:56-113): CARLA waypoints + yaw up-sampling
// become a Philox4x32-10 sampler (or the same 4 m lattice x 7 yaws layout), and the O(W^2)
// Python double loop (cuda_agent.py:71-91) becomes a cell-list search that emits exactly the
// CSR arrays init_parameters['neighbors'] / ['num_neighbors'] hold.
//
// Neighbour rule (restated in oracle/world_oracle.py, the definition both sides follow):
// j is a neighbour of i iff (x_j, y_j) != (x_i, y_i) and
// (double)sqrtf(fl(fl(dx*dx) + fl(dy*dy))) <= disk_radius; rows list j ascending.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/gmt_c_api.h"



namespace {

// error text is owned by gmt_kernels.cu; these entry points report through their return code
#define RCK(call)                                  \
    do {                                           \
        cudaError_t e_ = (call);                   \
        if (e_ != cudaSuccess) {                   \
            fprintf(stderr, "gmt_roadmap: %s failed: %s\n", #call, cudaGetErrorString(e_)); \
            return GMT_ERR_CUDA;                   \
        }                                          \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., Random123): counter = (index, 0, 0, 0), key = (seed, 0x474D5421)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, unsigned k0, unsigned k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ float u01(unsigned r) { return __fmul_rn((float)(r >> 8), 5.9604644775390625e-8f); }

__global__ void sample_states_kernel(unsigned seed, int n, float side, int mode, float *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x, y, th;
    if (mode == 0) {
        const uint4 r = philox4x32_10(make_uint4((unsigned)i, 0u, 0u, 0u), seed, 0x474D5421u);
        x = __fmul_rn(u01(r.x), side);
        y = __fmul_rn(u01(r.y), side);
        th = __fsub_rn(__fmul_rn(u01(r.z), 6.2831855f), 3.1415927f);
    } else {
        // 4 m lattice, 7 yaws per location: k*45 deg for k != 4, wrapped as cuda_agent.py:102-107
        const int W = (int)floor((double)side / 4.0) + 1;
        const int loc = i / 7, k7 = i % 7;
        const int k = k7 >= 4 ? k7 + 1 : k7;
        double theta = (double)k * 360.0 / 8.0;
        if (theta >= 180.0) theta -= 360.0;
        x = (float)((double)(loc % W) * 4.0);
        y = (float)((double)(loc / W) * 4.0);
        th = (float)(theta * 3.141592653589793 / 180.0);
    }
    out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = th;
}

// ---------------------------------------------------------------------------------------------
// cell grid
// ---------------------------------------------------------------------------------------------
struct CellGrid { float x0, y0, inv_h; int nx, ny; };

__device__ __forceinline__ int cell_of(const CellGrid &g, float x, float y, int &cx, int &cy)
{
    cx = min(max(__float2int_rd(__fmul_rn(__fsub_rn(x, g.x0), g.inv_h)), 0), g.nx - 1);
    cy = min(max(__float2int_rd(__fmul_rn(__fsub_rn(y, g.y0), g.inv_h)), 0), g.ny - 1);
    return cy * g.nx + cx;
}

__global__ void cell_count_kernel(const float *__restrict__ xyt, int n, CellGrid g, int *cell_of_state, int *cell_cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx, cy;
    const int c = cell_of(g, xyt[3 * i], xyt[3 * i + 1], cx, cy);
    cell_of_state[i] = c;
    atomicAdd(&cell_cnt[c], 1);
}

// three-kernel exclusive scan (block sums -> scan of sums -> add back); in place on `a`, a[n] = total
#define SCAN_BLOCK 1024
__global__ void scan_block_kernel(int *a, int n, int *block_sums)
{
    __shared__ int s[SCAN_BLOCK];
    const int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const int v = i < n ? a[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
        const int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n) a[i] = s[threadIdx.x] - v;
    if (threadIdx.x == SCAN_BLOCK - 1) block_sums[blockIdx.x] = s[threadIdx.x];
}
__global__ void scan_sums_kernel(int *block_sums, int nb, long long *total)
{
    // single thread block, serial over chunks: nb is n / 1024
    __shared__ int s[SCAN_BLOCK];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += SCAN_BLOCK) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? block_sums[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
            const int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nb) block_sums[i] = (int)(carry + s[threadIdx.x] - v);
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1) carry += s[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void scan_add_kernel(int *a, int n, const int *block_sums, const long long *total)
{
    const int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (i < n) a[i] += block_sums[blockIdx.x];
    if (i == 0) a[n] = (int)*total;
}

__global__ void cell_scatter_kernel(const float *__restrict__ xyt, int n, const int *__restrict__ cell_of_state,
                                    int *cell_fill, float4 *sorted)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pos = atomicAdd(&cell_fill[cell_of_state[i]], 1);
    sorted[pos] = make_float4(xyt[3 * i], xyt[3 * i + 1], __int_as_float(i), 0.f);   // SoA-style float4 record
}

// The rule is (double)sqrtf(q) <= r with q = fl(fl(dx*dx) + fl(dy*dy)).  sqrtf is monotone, so it is the same as
// q <= qmax where qmax is the LARGEST float whose square root passes (found on the host by bisection over the float
// bit patterns, q_threshold below): one compare instead of a square root, a conversion and a double compare.
__device__ __forceinline__ bool is_neighbour(float xi, float yi, float4 c, float qmax)
{
    if (c.x == xi && c.y == yi) return false;                     // same location (cuda_agent.py:78)
    const float dx = __fsub_rn(xi, c.x), dy = __fsub_rn(yi, c.y);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) <= qmax;   // inclusive (cuda_agent.py:80)
}

// bitonic sort of up to 64 ints held two per lane (slot = lane + 32 * half), ascending; pads are INT_MAX
__device__ __forceinline__ void warp_sort64(int &a, int &b, int lane)
{
#pragma unroll
    for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j >= 1; j >>= 1) {
            if (j == 32) {                                        // partner = the other half, same lane (k = 64: ascending)
                const int lo = min(a, b), hi = max(a, b);
                a = lo; b = hi;
            } else {
                const int pa = __shfl_xor_sync(0xffffffffu, a, j), pb = __shfl_xor_sync(0xffffffffu, b, j);
                const bool lower = (lane & j) == 0;               // I hold the lower slot of the pair
                const bool up_a = (lane & k) == 0 || k == 64;     // slots 0..31: bit k of the slot = bit k of the lane
                const bool up_b = ((lane + 32) & k) == 0;         // slots 32..63
                a = (lower == up_a) ? min(a, pa) : max(a, pa);
                b = (lower == up_b) ? min(b, pb) : max(b, pb);
            }
        }
    }
}

// pass 0: count row lengths; pass 1: write the rows, ascending.  One WARP per state, states taken in cell order
// (neighbouring warps scan the same 3 x 3 cells: L1/L2 hits); the members of a cell row are contiguous float4
// records, so the 32 lanes stream them coalesced and test 32 candidates at once.  Pass 1 compacts the row into
// shared memory by ballot, rank-sorts it there (ids are distinct: rank = number of smaller ids) and writes it
// out in order.  Rows longer than ROW_SMEM fall back to repeated warp-wide minimum selection (slow, never seen
// on the BASELINE maps where rows hold ~55 entries).
#define NBR_WARPS 8
#define ROW_SMEM 256
template <int kPass>
__global__ void __launch_bounds__(NBR_WARPS * 32) neighbour_kernel(int n, CellGrid g, float r,
                                 const int *__restrict__ cell_start, const float4 *__restrict__ sorted,
                                 int *row_cnt_or_off, int *vals)
{
    __shared__ int s_row[kPass ? NBR_WARPS : 1][kPass ? ROW_SMEM : 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int s = blockIdx.x * NBR_WARPS + wib;
    if (s >= n) return;
    const float4 me = sorted[s];
    const float xi = me.x, yi = me.y;
    const int i = __float_as_int(me.z);
    int cx, cy;
    cell_of(g, xi, yi, cx, cy);
    int cnt = 0;
    for (int yy = max(cy - 1, 0); yy <= min(cy + 1, g.ny - 1); yy++) {
        const int ca = yy * g.nx + max(cx - 1, 0), cb = yy * g.nx + min(cx + 1, g.nx - 1);
        const int e = cell_start[cb + 1];
        for (int kb = cell_start[ca]; kb < e; kb += 32) {         // three cells of a row are contiguous
            const int k = kb + lane;
            float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
            bool hit = false;
            if (k < e) { c = sorted[k]; hit = is_neighbour(xi, yi, c, r); }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (kPass && hit) {
                const int pos = cnt + __popc(m & ((1u << lane) - 1u));
                if (pos < ROW_SMEM) s_row[wib][pos] = __float_as_int(c.z);
            }
            cnt += __popc(m);
        }
    }
    if (!kPass) { if (lane == 0) row_cnt_or_off[i] = cnt; return; }
    int *row = vals + row_cnt_or_off[i];
    __syncwarp();
    if (cnt <= 64) {                                               // the usual row (~55 entries): sort in registers
        int a = lane < cnt ? s_row[wib][lane] : 0x7fffffff, b = lane + 32 < cnt ? s_row[wib][lane + 32] : 0x7fffffff;
        warp_sort64(a, b, lane);
        if (lane < cnt) row[lane] = a;
        if (lane + 32 < cnt) row[lane + 32] = b;
        return;
    }
    if (cnt <= ROW_SMEM) {
        for (int a = lane; a < cnt; a += 32) {
            const int v = s_row[wib][a];
            int rank = 0;
            for (int b = 0; b < cnt; b++) rank += s_row[wib][b] < v;      // broadcast reads
            row[rank] = v;
        }
        return;
    }
    int last = -1;                                                 // generic fallback: next smallest id, cnt times
    for (int out = 0; out < cnt; out++) {
        int best = 0x7fffffff;
        for (int yy = max(cy - 1, 0); yy <= min(cy + 1, g.ny - 1); yy++) {
            const int ca = yy * g.nx + max(cx - 1, 0), cb = yy * g.nx + min(cx + 1, g.nx - 1);
            const int e = cell_start[cb + 1];
            for (int k = cell_start[ca] + lane; k < e; k += 32) {
                const float4 c = sorted[k];
                const int id = __float_as_int(c.z);
                if (id > last && id < best && is_neighbour(xi, yi, c, r)) best = id;
            }
        }
        best = __reduce_min_sync(0xffffffffu, best);
        if (lane == 0) row[out] = best;
        last = best;
    }
}

}  // namespace

extern "C" {

int gmt_sample_states(unsigned seed, int n, float side, int mode, int device, float *out_states_xyt)
{
    if (n < 0 || !out_states_xyt || (mode != 0 && mode != 1) || !(side > 0.f)) return GMT_ERR_INVALID;
    if (n == 0) return GMT_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return GMT_ERR_CUDA;
    RCK(cudaSetDevice(device));
    float *d = nullptr;
    RCK(cudaMalloc(&d, sizeof(float) * 3 * (size_t)n));
    sample_states_kernel<<<(n + 255) / 256, 256>>>(seed, n, side, mode, d);
    RCK(cudaGetLastError());
    RCK(cudaMemcpy(out_states_xyt, d, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return GMT_OK;
}

// largest float q with (double)sqrtf(q) <= r (r >= 0): bisection over the bit patterns of the non-negative floats
static float q_threshold(double r)
{
    unsigned lo = 0u, hi = 0x7f7fffffu;                           // invariant: pattern lo passes
    if ((double)sqrtf(0.0f) > r) return -1.0f;                    // (r < 0 is rejected by the caller)
    while (lo < hi) {
        const unsigned mid = lo + (hi - lo + 1u) / 2u;
        float q; memcpy(&q, &mid, 4);
        if ((double)sqrtf(q) <= r) lo = mid; else hi = mid - 1u;
    }
    float q; memcpy(&q, &lo, 4);
    return q;
}

}  // extern "C"

// The search itself, device to device: d_xyt [3n] in, *d_off [n_alloc + 1] (exclusive row offsets, *d_off[n] = total)
// and *d_vals out (cudaMalloc'ed here; the caller frees them).  Bounding box of the states from the host copy.
int gmt_roadmap_build_csr_device(const float *states_xyt_host, const float *d_xyt, int n, int n_alloc, double disk_radius,
                                 int **d_off_out, int **d_vals_out, int **d_counts_out, long long *total)
{
    float x0 = INFINITY, y0 = INFINITY, x1 = -INFINITY, y1 = -INFINITY;
    for (int i = 0; i < n; i++) {
        const float x = states_xyt_host[3 * i], y = states_xyt_host[3 * i + 1];
        if (!(fabsf(x) < 1e30f) || !(fabsf(y) < 1e30f)) return GMT_ERR_INVALID;
        x0 = std::min(x0, x); y0 = std::min(y0, y); x1 = std::max(x1, x); y1 = std::max(y1, y);
    }
    // cell edge a little above the radius so that every pair within the radius sits in adjacent cells
    float h = (float)(disk_radius * 1.001 + 1e-3);
    const float ext = std::max(x1 - x0, y1 - y0);
    h = std::max(h, ext / 4000.0f);
    CellGrid g;
    g.x0 = x0; g.y0 = y0; g.inv_h = 1.0f / h;
    g.nx = (int)floorf((x1 - x0) * g.inv_h) + 2;
    g.ny = (int)floorf((y1 - y0) * g.inv_h) + 2;
    const int ncell = g.nx * g.ny;

    const size_t N = (size_t)n;
    int *d_cell = nullptr, *d_start = nullptr, *d_fill = nullptr, *d_sums = nullptr;
    int *d_row = nullptr, *d_vals = nullptr, *d_counts = nullptr; float4 *d_sorted = nullptr; long long *d_total = nullptr;
    const int nbc = (ncell + SCAN_BLOCK - 1) / SCAN_BLOCK, nbr = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    auto release = [&](bool all) {
        cudaFree(d_cell); cudaFree(d_start); cudaFree(d_fill); cudaFree(d_sums); cudaFree(d_sorted); cudaFree(d_total);
        if (all) { cudaFree(d_row); cudaFree(d_vals); cudaFree(d_counts); }
    };
#define BCK(call)                                  \
    do {                                           \
        cudaError_t e_ = (call);                   \
        if (e_ != cudaSuccess) {                   \
            fprintf(stderr, "gmt_roadmap: %s failed: %s\n", #call, cudaGetErrorString(e_)); \
            release(true);                         \
            return e_ == cudaErrorMemoryAllocation ? GMT_ERR_NOMEM : GMT_ERR_CUDA; \
        }                                          \
    } while (0)
    BCK(cudaMalloc(&d_cell, sizeof(int) * N));
    BCK(cudaMalloc(&d_start, sizeof(int) * ((size_t)ncell + 1)));
    BCK(cudaMalloc(&d_fill, sizeof(int) * ((size_t)ncell + 1)));
    BCK(cudaMalloc(&d_sums, sizeof(int) * (size_t)std::max(nbc, nbr)));
    BCK(cudaMalloc(&d_row, sizeof(int) * ((size_t)std::max(n_alloc, n) + 1)));
    BCK(cudaMalloc(&d_sorted, sizeof(float4) * N));
    BCK(cudaMalloc(&d_total, sizeof(long long)));
    BCK(cudaMemset(d_start, 0, sizeof(int) * ((size_t)ncell + 1)));

    const int tb = 256, gb = (n + tb - 1) / tb;
    cell_count_kernel<<<gb, tb>>>(d_xyt, n, g, d_cell, d_start);
    scan_block_kernel<<<nbc, SCAN_BLOCK>>>(d_start, ncell, d_sums);
    scan_sums_kernel<<<1, SCAN_BLOCK>>>(d_sums, nbc, d_total);
    scan_add_kernel<<<nbc, SCAN_BLOCK>>>(d_start, ncell, d_sums, d_total);
    BCK(cudaMemcpy(d_fill, d_start, sizeof(int) * ((size_t)ncell + 1), cudaMemcpyDeviceToDevice));
    cell_scatter_kernel<<<gb, tb>>>(d_xyt, n, d_cell, d_fill, d_sorted);
    const int gw = (n + NBR_WARPS - 1) / NBR_WARPS;
    const float qmax = q_threshold(disk_radius);
    neighbour_kernel<0><<<gw, NBR_WARPS * 32>>>(n, g, qmax, d_start, d_sorted, d_row, nullptr);
    BCK(cudaGetLastError());
    if (d_counts_out) {                                        // row lengths, before the scan overwrites them
        BCK(cudaMalloc(&d_counts, sizeof(int) * N));
        BCK(cudaMemcpy(d_counts, d_row, sizeof(int) * N, cudaMemcpyDeviceToDevice));
    }
    scan_block_kernel<<<nbr, SCAN_BLOCK>>>(d_row, n, d_sums);
    scan_sums_kernel<<<1, SCAN_BLOCK>>>(d_sums, nbr, d_total);
    scan_add_kernel<<<nbr, SCAN_BLOCK>>>(d_row, n, d_sums, d_total);
    long long tot = 0;
    BCK(cudaMemcpy(&tot, d_total, sizeof(long long), cudaMemcpyDeviceToHost));
    if (tot > 0x7fffffffLL) { release(true); return GMT_ERR_INVALID; }
    BCK(cudaMalloc(&d_vals, sizeof(int) * (size_t)std::max<long long>(tot, 1)));
    neighbour_kernel<1><<<gw, NBR_WARPS * 32>>>(n, g, qmax, d_start, d_sorted, d_row, d_vals);
    BCK(cudaGetLastError());
    BCK(cudaDeviceSynchronize());
#undef BCK
    release(false);
    *d_off_out = d_row; *d_vals_out = d_vals;
    if (d_counts_out) *d_counts_out = d_counts;
    if (total) *total = tot;
    return GMT_OK;
}

extern "C" {

int gmt_build_neighbors(const float *states_xyt, int n, double disk_radius, int device, int **vals, int **counts,
                        long long *total)
{
    if (!states_xyt || n <= 0 || !vals || !counts || !(disk_radius >= 0.0)) return GMT_ERR_INVALID;
    *vals = nullptr; *counts = nullptr;
    if (total) *total = 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return GMT_ERR_CUDA;
    RCK(cudaSetDevice(device));
    const size_t N = (size_t)n;
    float *d_xyt = nullptr;
    RCK(cudaMalloc(&d_xyt, sizeof(float) * 3 * N));
    if (cudaMemcpy(d_xyt, states_xyt, sizeof(float) * 3 * N, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d_xyt); return GMT_ERR_CUDA; }
    int *d_off = nullptr, *d_vals = nullptr, *d_counts = nullptr;
    long long tot = 0;
    const int rc = gmt_roadmap_build_csr_device(states_xyt, d_xyt, n, n, disk_radius, &d_off, &d_vals, &d_counts, &tot);
    cudaFree(d_xyt);
    if (rc != GMT_OK) return rc;
    int *h_counts = (int *)malloc(sizeof(int) * N);
    int *h_vals = (int *)malloc(sizeof(int) * (size_t)std::max<long long>(tot, 1));
    bool ok = h_counts && h_vals &&
              cudaMemcpy(h_counts, d_counts, sizeof(int) * N, cudaMemcpyDeviceToHost) == cudaSuccess &&
              cudaMemcpy(h_vals, d_vals, sizeof(int) * (size_t)tot, cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(d_off); cudaFree(d_vals); cudaFree(d_counts);
    if (!ok) { free(h_counts); free(h_vals); return (h_counts && h_vals) ? GMT_ERR_CUDA : GMT_ERR_NOMEM; }
    *vals = h_vals; *counts = h_counts;
    if (total) *total = tot;
    return GMT_OK;
}

}  // extern "C"
