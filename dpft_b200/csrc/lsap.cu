// Hungarian assignment on the GPU: one warp per sample solves the rectangular linear-sum-assignment
// problem of the reference's HungarianAnassigner (src/dprt/training/assigner.py:134-141: `C.cpu()` + scipy's
// linear_sum_assignment per sample) on the device, so the criterion needs no host synchronisation and the whole training
// step stays capturable in one CUDA graph.  The algorithm is in lsap_core.h (shared with the host harness that checks it
// against scipy; on B200: tests/test_criterion_metrics_gpu.py).  dpft_b200.criterion uses it on request (lsap_solver="device").
#include "common.cuh"
#include "lsap_core.h"

namespace dpft {
namespace {

__global__ void __launch_bounds__(32)
lsap_kernel(const float* __restrict__ cost, const int* __restrict__ counts, long long* __restrict__ col4row, int N, int Mmax) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, lane = threadIdx.x;
    __shared__ lsap::Workspace w;
    if (lane == 0) {
        double* d = reinterpret_cast<double*>(smem_raw);
        w.v = d;
        w.shortest = d + N;
        w.path = reinterpret_cast<int*>(d + 2 * N);
        w.row4col = w.path + N;
        w.SC = reinterpret_cast<unsigned char*>(w.row4col + N);
    }
    __syncwarp();
    int R = counts[b];
    R = R < 0 ? 0 : (R > Mmax ? Mmax : R);
    int st = 0;
    if (R > 0) st = lsap::solve<32>(cost + (long long)b * N * Mmax, Mmax, R, N, w, lane);
    __syncwarp();
    for (int i = lane; i < Mmax; i += 32) col4row[(long long)b * Mmax + i] = (i < R && st == 0) ? (long long)w.col4row[i] : -1LL;
}

}  // namespace
}  // namespace dpft

using namespace dpft;

extern "C" int dpft_lsap_forward(const float* cost, const int* counts, long long* col4row, int B, int N, int Mmax, void* stream) {
    DPFT_REQUIRE(cost && counts && col4row, "lsap: null pointer");
    DPFT_REQUIRE(B >= 1 && N >= 1 && Mmax >= 1, "lsap: bad sizes B=%d N=%d Mmax=%d", B, N, Mmax);
    DPFT_REQUIRE(Mmax <= lsap::kMaxRows, "lsap: at most %d ground-truth boxes per sample (got %d)", lsap::kMaxRows, Mmax);
    DPFT_REQUIRE(Mmax <= N, "lsap: more rows (%d) than columns (%d)", Mmax, N);
    const size_t smem = (size_t)N * (2 * sizeof(double) + 2 * sizeof(int) + 1) + 16;
    DPFT_REQUIRE(smem <= 200 * 1024, "lsap: N=%d columns need %zu bytes of shared memory", N, smem);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        int st = cuda_status(cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "lsap attr");
        if (st) return st;
        configured = smem;
    }
    lsap_kernel<<<B, 32, smem, (cudaStream_t)stream>>>(cost, counts, col4row, N, Mmax);
    DPFT_LAUNCH_CHECK("lsap_kernel");
    return DPFT_OK;
}
