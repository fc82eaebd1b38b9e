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

// Candidate order of the column search: smaller reduced cost first; ties: an unassigned column wins, then the smaller index.
// One rule for the scan inside a lane and for the reduction across lanes, so every lane ends with the same winner.
struct Candidate {
    double val;
    int idx;        // -1 = none
    int free_col;   // 1 = the column is unassigned
};
LSAP_HD bool better(const Candidate& a, const Candidate& b) {          // a beats b
    if (a.idx < 0) return false;
    if (b.idx < 0) return true;
    if (a.val != b.val) return a.val < b.val;
    if (a.free_col != b.free_col) return a.free_col > b.free_col;
    return a.idx < b.idx;
}

template <int LANES> LSAP_HD void reduce_min(Candidate& c) {
#if defined(__CUDA_ARCH__)
    if (LANES > 1) {
        for (int o = LANES / 2; o > 0; o >>= 1) {
            Candidate other;
            other.val = __shfl_xor_sync(0xffffffffu, c.val, o);
            other.idx = __shfl_xor_sync(0xffffffffu, c.idx, o);
            other.free_col = __shfl_xor_sync(0xffffffffu, c.free_col, o);
            if (better(other, c)) c = other;
        }
    }
#endif
}

// The pieces of one augmentation step; `lane` of `lanes` walks its share of the rows / columns.
LSAP_HD void init_problem(int R, int C, Workspace& w, int lane, int lanes) {
    for (int j = lane; j < C; j += lanes) { w.v[j] = 0.0; w.row4col[j] = -1; }
    for (int i = lane; i < R; i += lanes) { w.u[i] = 0.0; w.col4row[i] = -1; }
}
LSAP_HD void init_search(int R, int C, Workspace& w, int lane, int lanes) {
    for (int j = lane; j < C; j += lanes) { w.shortest[j] = kInf; w.SC[j] = 0; }
    for (int i = lane; i < R; i += lanes) w.SR[i] = 0;
}
// relax the unscanned columns of this lane from row i and return the lane's best remaining column
// cost(r, c) = cost[c * ld + r]  (the (C, R)-shaped slice of the assigner's (B, N, Mmax) cost tensor: ld = Mmax)
LSAP_HD Candidate scan_lane(const float* cost, int ld, int C, int i, double min_val, Workspace& w, int lane, int lanes) {
    const double ui = w.u[i];
    Candidate best{kInf, -1, 0};
    for (int j = lane; j < C; j += lanes) {
        if (w.SC[j]) continue;
        const double r = min_val + (double)cost[(long long)j * ld + i] - ui - w.v[j];
        if (r < w.shortest[j]) { w.shortest[j] = r; w.path[j] = i; }
        const Candidate c{w.shortest[j], j, w.row4col[j] < 0 ? 1 : 0};
        if (better(c, best)) best = c;
    }
    return best;
}
LSAP_HD void update_duals(int R, int C, int cur, double min_val, Workspace& w, int lane, int lanes) {
    if (lane == 0) w.u[cur] += min_val;
    for (int r = lane; r < R; r += lanes)
        if (w.SR[r] && r != cur) w.u[r] += min_val - w.shortest[w.col4row[r]];
    for (int j = lane; j < C; j += lanes)
        if (w.SC[j]) w.v[j] -= min_val - w.shortest[j];
}
LSAP_HD void augment(int cur, int sink, Workspace& w) {               // along the path back to the current row
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

// Returns 0, or -1 when no finite-cost assignment exists.  col4row[r] = the column matched to row r.
template <int LANES>
LSAP_HD int solve(const float* cost, int ld, int R, int C, Workspace& w, int lane) {
    init_problem(R, C, w, lane, LANES);
    sync_lanes<LANES>();
    for (int cur = 0; cur < R; ++cur) {
        init_search(R, C, w, lane, LANES);
        sync_lanes<LANES>();
        double min_val = 0.0;
        int i = cur, sink = -1;
        while (sink < 0) {
            if (lane == 0) w.SR[i] = 1;
            Candidate best = scan_lane(cost, ld, C, i, min_val, w, lane, LANES);
            reduce_min<LANES>(best);
            if (best.idx < 0 || best.val >= kInf) return -1;
            min_val = best.val;
            const int j = best.idx;
            if (lane == 0) w.SC[j] = 1;
            sync_lanes<LANES>();
            if (w.row4col[j] < 0) sink = j;
            else i = w.row4col[j];
        }
        update_duals(R, C, cur, min_val, w, lane, LANES);
        sync_lanes<LANES>();
        if (lane == 0) augment(cur, sink, w);
        sync_lanes<LANES>();
    }
    return 0;
}

}  // namespace lsap
}  // namespace dpft
