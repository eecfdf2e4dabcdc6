// oracle/ref_gpu_harness.cuh -- TEST INFRASTRUCTURE ONLY.
//
// Appended (by oracle/build_oracle.py, in a temporary translation unit) after the
// reference's own CUDA source string (/root/reference/carla/kernel.py:13-493) so
// that nvcc compiles the reference's __device__ functions for sm_100a exactly as
// PyCuda's SourceModule would, plus this one extra kernel that exposes them per
// word.  Output: oracle/_ref/ref_kernel.cubin (git-ignored).  The GPU parity tests
// load it through cuda-python when it is present.
//
// Output format = gmto_edge_costs (oracle/gmt_oracle.c): cost4[i*4+w] word length
// (NaN when the word does not exist), verdict bit w = exists, bit 4+w = collides,
// best = computeDubinsCost's minimum over collision-free words (1e10f if none).
extern "C" __global__ void refh_edge_costs(float *states, const int *y, const int *x, int pairs,
                                           float radius, float *obstacles, int num_obs, float *cost4,
                                           unsigned *verdict_bits, float *best)
{
    int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i >= pairs) return;
    float *sp = &states[y[i] * 3];
    float *ep = &states[x[i] * 3];
    int r = (int)radius;
    unsigned bits = 0;
    for (int w = 0; w < 4; w++) {
        float fc = __int_as_float(0x7f800000), oc = __int_as_float(0x7f800000);
        if (w == 0) { RSRcost(&fc, sp, ep, r, obstacles, 0); RSRcost(&oc, sp, ep, r, obstacles, num_obs); }
        if (w == 1) { LSLcost(&fc, sp, ep, r, obstacles, 0); LSLcost(&oc, sp, ep, r, obstacles, num_obs); }
        if (w == 2) { LSRcost(&fc, sp, ep, r, obstacles, 0); LSRcost(&oc, sp, ep, r, obstacles, num_obs); }
        if (w == 3) { RSLcost(&fc, sp, ep, r, obstacles, 0); RSLcost(&oc, sp, ep, r, obstacles, num_obs); }
        bool exists = fc != __int_as_float(0x7f800000);
        cost4[i * 4 + w] = exists ? fc : __int_as_float(0x7fc00000);
        if (exists) bits |= 1u << w;
        if (exists && oc == __int_as_float(0x7f800000)) bits |= 1u << (4 + w);
    }
    verdict_bits[i] = bits;
    float cst = __int_as_float(0x7f800000), zero = 0.0f;
    computeDubinsCost(cst, zero, ep, sp, radius, obstacles, num_obs);
    best[i] = cst;
}
