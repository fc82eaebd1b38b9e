// Rectangular linear-sum-assignment (minimum-cost matching of every one of R rows to a distinct one of C >= R columns) by
// shortest augmenting paths with dual variables — the algorithm scipy.optimize.linear_sum_assignment implements (Crouse, "On
// implementing 2D rectangular assignment algorithms", 2016), which the reference calls per sample on the CPU for its
// Hungarian assigner (src/dprt/training/assigner.py:136).  Rows = ground-truth boxes (a few), columns = queries (hundreds).
//
// One source for two builds: LANES = 32 inside lsap_kernel (one warp per sample: every scan over the columns is strided
// over the lanes and finished with a shuffle reduction; all lanes follow the same control flow), LANES = 1 in the host
// harness (tools/lsap_host.cpp), where the identical code runs sequentially and is checked against scipy on the CPU
// (tests/test_lsap.py) — there is no GPU in the build container.  Arithmetic in double, as scipy's.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LSAP_HD __host__ __device__ __forceinline__
#else
#define LSAP_HD inline
#endif

namespace dpft {
namespace lsap {

constexpr int kMaxRows = 64;
constexpr double kInf = 1e300;

struct Workspace {          // per problem; in shared memory on the device
    double* v;              // [C] column duals
    double* shortest;       // [C]
    int* path;              // [C] predecessor row of a column on the current shortest-path tree
    int* row4col;           // [C] assignment, -1 = free
    unsigned char* SC;      // [C] column scanned
    double u[kMaxRows];     // row duals
    int col4row[kMaxRows];
    unsigned char SR[kMaxRows];
};

template <int LANES> LSAP_HD void sync_lanes() {
#if defined(__CUDA_ARCH__)
    if (LANES > 1) __syncwarp();
#endif
}

// (value, index) minimum over the lanes; ties: an unassigned column wins, then the smaller index — on every lane
template <int LANES> LSAP_HD void reduce_min(double& val, int& idx, int& free_col) {
#if defined(__CUDA_ARCH__)
    if (LANES > 1) {
        for (int o = LANES / 2; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, val, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            const int of = __shfl_xor_sync(0xffffffffu, free_col, o);
            const bool better = oi >= 0 && (idx < 0 || ov < val || (ov == val && (of > free_col || (of == free_col && oi < idx))));
            if (better) { val = ov; idx = oi; free_col = of; }
        }
    }
#endif
}

// cost(r, c) = cost[c * ld + r]  (the (C, R)-shaped slice of the assigner's (B, N, Mmax) cost tensor: ld = Mmax).
// Returns 0, or -1 when no finite-cost assignment exists.  col4row[r] = the column matched to row r.
template <int LANES>
LSAP_HD int solve(const float* cost, int ld, int R, int C, Workspace& w, int lane) {
    for (int j = lane; j < C; j += LANES) { w.v[j] = 0.0; w.row4col[j] = -1; }
    for (int i = lane; i < R; i += LANES) { w.u[i] = 0.0; w.col4row[i] = -1; }
    sync_lanes<LANES>();
    for (int cur = 0; cur < R; ++cur) {
        for (int j = lane; j < C; j += LANES) { w.shortest[j] = kInf; w.SC[j] = 0; }
        for (int i = lane; i < R; i += LANES) w.SR[i] = 0;
        sync_lanes<LANES>();
        double min_val = 0.0;
        int i = cur, sink = -1;
        while (sink < 0) {
            if (lane == 0) w.SR[i] = 1;
            const double ui = w.u[i];
            double best = kInf;
            int best_j = -1, best_free = 0;
            for (int j = lane; j < C; j += LANES) {
                if (w.SC[j]) continue;
                const double r = min_val + (double)cost[(long long)j * ld + i] - ui - w.v[j];
                if (r < w.shortest[j]) { w.shortest[j] = r; w.path[j] = i; }
                const double s = w.shortest[j];
                const int fr = w.row4col[j] < 0 ? 1 : 0;
                if (best_j < 0 || s < best || (s == best && fr > best_free)) { best = s; best_j = j; best_free = fr; }
            }
            reduce_min<LANES>(best, best_j, best_free);
            if (best_j < 0 || best >= kInf) return -1;
            min_val = best;
            const int j = best_j;
            if (lane == 0) w.SC[j] = 1;
            sync_lanes<LANES>();
            if (w.row4col[j] < 0) sink = j;
            else i = w.row4col[j];
        }
        // dual updates (rows / columns of the alternating tree)
        if (lane == 0) w.u[cur] += min_val;
        for (int r = lane; r < R; r += LANES)
            if (w.SR[r] && r != cur) w.u[r] += min_val - w.shortest[w.col4row[r]];
        for (int j = lane; j < C; j += LANES)
            if (w.SC[j]) w.v[j] -= min_val - w.shortest[j];
        sync_lanes<LANES>();
        // augment along the path back to the current row
        if (lane == 0) {
            int j = sink;
            while (true) {
                const int r = w.path[j];
                w.row4col[j] = r;
                const int prev = w.col4row[r];
                w.col4row[r] = j;
                j = prev;
                if (r == cur) break;
            }
        }
        sync_lanes<LANES>();
    }
    return 0;
}

}  // namespace lsap
}  // namespace dpft
