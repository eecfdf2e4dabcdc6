// oracle/ref_host_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the reference's OWN kernel source for the host CPU.  The CUDA-C
// string in /root/reference/carla/kernel.py:13-493 is extracted at build time
// by oracle/build_oracle.py into a temporary file (never into this repo) and
// pulled in below, unmodified, through REF_SOURCE_FILE.  This file only adds
// what PyCuda/nvcc would otherwise provide (__device__/__global__, the thread
// index built-ins) and C entry points that run each kernel over an emulated
// grid.  Output: oracle/_ref/libref_host.so (git-ignored).
//
// Used (a) to pin oracle/gmt_oracle.c (libm flavour) bit for bit and (b) as the
// "reference" CPU baseline of bench.py (--impl reference / cpu_baseline).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <cmath>
#include <cstdlib>

#ifdef _OPENMP
#include <omp.h>
#endif

using std::abs;
using std::isnan;

#define __device__
#define __global__

struct ref_dim3 { unsigned x, y, z; };
static thread_local ref_dim3 threadIdx, blockIdx, blockDim;

extern "C" {
#include REF_SOURCE_FILE
}

typedef void (*word_fn)(float *, float *, float *, int, float *, int);
static word_fn kWords[4] = {RSRcost, LSLcost, LSRcost, RSLcost};  // kernel.py:421-424 order

static inline void set_thread(int global_index, int tpb)
{
    blockDim.x = (unsigned)tpb; blockDim.y = blockDim.z = 1;
    blockIdx.x = (unsigned)(global_index / tpb); blockIdx.y = blockIdx.z = 0;
    threadIdx.x = (unsigned)(global_index % tpb); threadIdx.y = threadIdx.z = 0;
}

extern "C" {

__attribute__((visibility("default"))) int ref_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// bench.py's reference arm runs under torchrun, which exports OMP_NUM_THREADS=1: the arm asks for all cores itself
__attribute__((visibility("default"))) void ref_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}

// dubinConnection over ceil(xSize/tpb) blocks of tpb threads (gmt_planner.py:200-201).
__attribute__((visibility("default"))) void ref_dubin_connection(
    float *cost, int *parent, int *x, int *y, float *states, int *open, int *unexplored, int xSize,
    int ySize, float *obstacles, int num_obs, float radius, int tpb)
{
    int nthreads = ((xSize + tpb - 1) / tpb) * tpb;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < nthreads; t++) {
        set_thread(t, tpb);
        int xs = xSize, ys = ySize, no = num_obs;
        float r = radius;
        dubinConnection(cost, parent, x, y, states, open, unexplored, &xs, &ys, obstacles, &no, &r);
    }
}

__attribute__((visibility("default"))) void ref_wavefront(int *G, int *open, float *cost,
                                                          float *threshold, int n, int tpb)
{
    int nthreads = ((n + tpb - 1) / tpb) * tpb;
    for (int t = 0; t < nthreads; t++) { set_thread(t, tpb); wavefront(G, open, cost, threshold, &n); }
}

__attribute__((visibility("default"))) void ref_grow_threshold(float *threshold, float *amount)
{
    set_thread(0, 1);
    growThreshold(threshold, amount);
}

__attribute__((visibility("default"))) void ref_neighbor_indicator(
    int *x_indicator, int *G, int *unexplored, int *neighbors, int *num_neighbors,
    int *neighbors_index, int gSize, int tpb)
{
    int nthreads = ((gSize + tpb - 1) / tpb) * tpb;
    for (int t = 0; t < nthreads; t++) {
        set_thread(t, tpb);
        neighborIndicator(x_indicator, G, unexplored, neighbors, num_neighbors, neighbors_index, &gSize);
    }
}

__attribute__((visibility("default"))) void ref_compact(int *x, int *scan, int *indicator,
                                                        int *waypoints, int n, int tpb)
{
    int nthreads = ((n + tpb - 1) / tpb) * tpb;
    for (int t = 0; t < nthreads; t++) { set_thread(t, tpb); compact(x, scan, indicator, waypoints, &n); }
}

__attribute__((visibility("default"))) int ref_compute_dubins_cost(
    float *cost, float parentCost, float *end_point, float *start_point, float r_min,
    float *obstacles, int num_obs)
{
    return computeDubinsCost(*cost, parentCost, end_point, start_point, r_min, obstacles, num_obs) ? 1 : 0;
}

// Classify one word by calling the reference function twice (without / with obstacles).
// flags: bit0 = word exists (finite cost written), bit1 = it collides.  *cost_out = its length.
__attribute__((visibility("default"))) int ref_word(int word, float *start_point, float *end_point,
                                                    int r_min, float *obstacles, int num_obs,
                                                    float *cost_out)
{
    float free_cost = INFINITY;
    kWords[word](&free_cost, start_point, end_point, r_min, obstacles, 0);
    if (cost_out) *cost_out = (free_cost == INFINITY) ? NAN : free_cost;
    if (free_cost == INFINITY) return 0;
    float c = INFINITY;
    kWords[word](&c, start_point, end_point, r_min, obstacles, num_obs);
    return 1 | ((c == INFINITY) ? 2 : 0);
}

// Same output format as gmto_edge_costs (oracle/gmt_oracle.c).
__attribute__((visibility("default"))) void ref_edge_costs(
    float *states, const int *y, const int *x, long pairs, float radius, float *obstacles,
    int num_obs, float *cost4, unsigned *verdict_bits, float *best)
{
    const int r = (int)radius;
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < pairs; i++) {
        unsigned bits = 0;
        for (int w = 0; w < 4; w++) {
            float wc;
            int fl = ref_word(w, &states[y[i] * 3], &states[x[i] * 3], r, obstacles, num_obs, &wc);
            if (cost4) cost4[i * 4 + w] = wc;
            if (fl & 1) bits |= 1u << w;
            if (fl & 2) bits |= 1u << (4 + w);
        }
        if (verdict_bits) verdict_bits[i] = bits;
        if (best) {
            // computeDubinsCost with cost = +inf and parentCost = 0 leaves min-over-free-words (or 1e10f)
            float cst = INFINITY, zero = 0.0f;
            computeDubinsCost(cst, zero, &states[x[i] * 3], &states[y[i] * 3], radius, obstacles, num_obs);
            best[i] = cst;
        }
    }
}

}  // extern "C"
