// match_grid.cu — StVO::matchGrid, points and lines (src/matching.cpp:111-177, :179-258): the stereo (left/right)
// matching step one row upstream of the hot path.  Same 256-bit Hamming primitive, candidates
// restricted to a window of the 64 x 48 bucket grid.
//
// The reference is sequential in the query index because of the bestLRMatches gate (:145-150): a train is considered by
// query i1 only if its distance beats the train's running minimum over all EARLIER queries.  That is a strict
// prefix-minimum per train over ascending query index, so it parallelises as
//   A. one thread per query: enumerate the window's candidates from the train grid (de-duplicated, direction-filtered for
//      lines), compute distances, emit (i2, d) pairs;
//   B. counting sort of the pairs by train;
//   C. one thread per train: a pair is "seen" iff no pair of the same train with a smaller query index has a distance
//      <= its own; the last record-setter is matches_21[i2];
//   D. one thread per query: best / second best over its seen pairs (with multiplicity), ratio test in double, then the
//      mutual filter (:166-174).
// One CTA per frame; a batch of frames runs in parallel.  Integer work and a handful of doubles only.
#include "common.cuh"

namespace plstvo {

namespace {

constexpr int MG_CAP = 128;        // candidates per query window; pair keys pack (query:16 | slot:7 | distance:9)
constexpr int MG_THREADS = 512;    // one CTA per frame: the work is a few thousand short dependent chains, so width buys latency

__device__ __forceinline__ int mg_distance(const uint8_t* a, const uint8_t* b) {   // StVO::distance (:93-109)
    const uint4* pa = reinterpret_cast<const uint4*>(a);
    const uint4* pb = reinterpret_cast<const uint4*>(b);
    const uint4 a0 = __ldg(pa), a1 = __ldg(pa + 1), b0 = __ldg(pb), b1 = __ldg(pb + 1);
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// LineIterator (src/lineIterator.cpp:34-77): calls f(x, y) for every visited cell
template <typename F>
__device__ __forceinline__ void mg_line_cells(double x1, double y1, double x2, double y2, F f) {
    const bool steep = fabs(y2 - y1) > fabs(x2 - x1);
    if (steep) { double t = x1; x1 = y1; y1 = t; t = x2; x2 = y2; y2 = t; }
    if (x1 > x2) { double t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
    const double dx = x2 - x1, dy = fabs(y2 - y1);
    double error = dx / 2.0;
    const int ystep = (y1 < y2) ? 1 : -1;
    int x = (int)x1, y = (int)y1;
    const int maxX = (int)x2;
    for (; !(x > maxX); ++x) {
        if (steep) f(y, x); else f(x, y);
        error -= dy;
        if (error < 0) { y += ystep; error += dx; }
    }
}

__device__ int mg_block_exclusive_scan(int* data, int n, int* s_tmp /* >= MG_THREADS ints */) {
    // in-place exclusive scan of data[0..n) by the whole block; returns the total
    const int tid = threadIdx.x;
    const int per = (n + MG_THREADS - 1) / MG_THREADS, lo = min(n, tid * per), hi = min(n, lo + per);
    int s = 0;
    for (int i = lo; i < hi; i++) s += data[i];
    s_tmp[tid] = s;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int i = 0; i < MG_THREADS; i++) { const int v = s_tmp[i]; s_tmp[i] = run; run += v; }
        s_tmp[MG_THREADS] = run;
    }
    __syncthreads();
    int run = s_tmp[tid];
    for (int i = lo; i < hi; i++) { const int v = data[i]; data[i] = run; run += v; }
    __syncthreads();
    return s_tmp[MG_THREADS];
}

template <bool LINES>
__global__ void __launch_bounds__(MG_THREADS) match_grid_kernel(const GridProblem* __restrict__ problems, GridParams prm) {
    extern __shared__ __align__(16) int smem_i[];
    int* cell_start = smem_i;                             // [cells + 1] train grid CSR
    int* cell_fill = smem_i + prm.rows * prm.cols + 1;    // [cells] counts, then fill cursors
    int* s_tmp = cell_fill + prm.rows * prm.cols;         // [MG_THREADS + 1]
    __shared__ int s_flag;
    const GridProblem pr = problems[blockIdx.x];
    const int tid = threadIdx.x, ncell = prm.rows * prm.cols, n1 = pr.n1, n2 = pr.n2;
    constexpr int CAP = MG_CAP;
    if (tid == 0) s_flag = 0;
    for (int q = tid; q < n1; q += MG_THREADS) pr.m12[q] = -1;
    if (n1 == 0 || n2 == 0) {
        if (tid == 0 && pr.count) *pr.count = 0;
        return;
    }

    // ---- train grid (stereoFrame.cpp:135-139 / :326-339): count, scan, fill ----
    for (int c = tid; c < ncell; c += MG_THREADS) cell_fill[c] = 0;
    __syncthreads();
    auto in_grid = [&](int x, int y) { return x >= 0 && x < prm.cols && y >= 0 && y < prm.rows; };
    for (int t = tid; t < n2; t += MG_THREADS) {
        if (!LINES) {
            const int x = pr.t_cell[2 * t], y = pr.t_cell[2 * t + 1];
            if (in_grid(x, y)) atomicAdd(&cell_fill[x * prm.rows + y], 1);
        } else {
            const double* L = pr.t_line + 4 * (size_t)t;
            mg_line_cells(L[0], L[1], L[2], L[3], [&](int x, int y) {
                if (in_grid(x, y)) atomicAdd(&cell_fill[x * prm.rows + y], 1);
            });
        }
    }
    __syncthreads();
    for (int c = tid; c < ncell; c += MG_THREADS) cell_start[c] = cell_fill[c];
    __syncthreads();
    const int n_items = mg_block_exclusive_scan(cell_start, ncell, s_tmp);
    if (tid == 0) cell_start[ncell] = n_items;
    for (int c = tid; c < ncell; c += MG_THREADS) cell_fill[c] = cell_start[c];
    __syncthreads();
    for (int t = tid; t < n2; t += MG_THREADS) {
        if (!LINES) {
            const int x = pr.t_cell[2 * t], y = pr.t_cell[2 * t + 1];
            if (in_grid(x, y)) pr.grid_items[atomicAdd(&cell_fill[x * prm.rows + y], 1)] = t;
        } else {
            const double* L = pr.t_line + 4 * (size_t)t;
            mg_line_cells(L[0], L[1], L[2], L[3], [&](int x, int y) {
                if (in_grid(x, y)) pr.grid_items[atomicAdd(&cell_fill[x * prm.rows + y], 1)] = t;
            });
        }
        pr.t_count[t] = 0;
    }
    __syncthreads();

    // ---- A. candidates of every query (:128-139 / :201-224) ----
    for (int q = tid; q < n1; q += MG_THREADS) {
        int2* seg = pr.q_pairs + (size_t)q * CAP;
        int k = 0;
        bool overflow = false;
        double vx = 0.0, vy = 0.0;
        const int* qc = pr.q_cell + (size_t)q * (LINES ? 4 : 2);
        if (LINES) {   // v = normalize(ep - sp) in INTEGER cell coordinates (:207-211): NaN when both share a cell
            vx = (double)(qc[2] - qc[0]);
            vy = (double)(qc[3] - qc[1]);
            const double mag = sqrt(vx * vx + vy * vy);
            vx /= mag;
            vy /= mag;
        }
        const uint8_t* dq = pr.d1 + (size_t)q * 32;
        for (int e = 0; e < (LINES ? 2 : 1); ++e) {   // GridStructure::get around sp (and ep)
            const int x = qc[2 * e], y = qc[2 * e + 1];
            const int min_x = max(0, x - prm.w.left), max_x = min(prm.cols, x + prm.w.right + 1);
            const int min_y = max(0, y - prm.w.up), max_y = min(prm.rows, y + prm.w.down + 1);
            for (int x_ = min_x; x_ < max_x; ++x_)
                for (int y_ = min_y; y_ < max_y; ++y_) {
                    const int c = x_ * prm.rows + y_;
                    for (int it = cell_start[c]; it < cell_start[c + 1]; ++it) {
                        const int i2 = pr.grid_items[it];
                        if (LINES) {
                            bool dup = false;      // the reference collects into an unordered_set
                            for (int j = 0; j < k; ++j) dup = dup || (seg[j].x == i2);
                            if (dup) continue;
                        }
                        if (k >= CAP) { overflow = true; continue; }
                        int d = -1;                // -1: in the candidate set but dropped by the direction filter
                        bool pass = true;
                        if (LINES) {
                            const double dt = vx * pr.t_dir[2 * i2] + vy * pr.t_dir[2 * i2 + 1];
                            if (fabs(dt) < prm.line_sim_th) pass = false;   // NaN passes (:221)
                        }
                        if (pass) d = mg_distance(dq, pr.d2 + (size_t)i2 * 32);
                        seg[k++] = make_int2(i2, d);
                        if (pass) atomicAdd(&pr.t_count[i2], 1);
                    }
                }
        }
        pr.q_count[q] = k;
        if (overflow) s_flag = 1;
    }
    __syncthreads();
    if (s_flag) {   // more than CAP candidates in one window: report, never guess
        if (tid == 0 && pr.count) *pr.count = PLSTVO_E_TOO_LARGE;
        return;
    }

    // ---- B. counting sort of the pairs by train ----
    for (int t = tid; t < n2; t += MG_THREADS) pr.t_start[t] = pr.t_count[t];
    __syncthreads();
    const int n_pairs = mg_block_exclusive_scan(pr.t_start, n2, s_tmp);
    if (tid == 0) pr.t_start[n2] = n_pairs;
    for (int t = tid; t < n2; t += MG_THREADS) pr.t_count[t] = pr.t_start[t];   // fill cursors
    __syncthreads();
    for (int q = tid; q < n1; q += MG_THREADS) {
        const int2* seg = pr.q_pairs + (size_t)q * CAP;
        const int k = pr.q_count[q];
        for (int j = 0; j < k; ++j)   // key = (query << 16) | (slot << 9) | distance: phase C never goes back to q_pairs
            if (seg[j].y >= 0) pr.t_slots[atomicAdd(&pr.t_count[seg[j].x], 1)] = (int)(((uint32_t)q << 16) | ((uint32_t)j << 9) | (uint32_t)seg[j].y);
    }
    __syncthreads();

    // ---- C. the gate (:145-150): strict prefix minimum over ascending query index, per train ----
    for (int t = tid; t < n2; t += MG_THREADS) {
        const int b = pr.t_start[t], e = pr.t_start[t + 1];
        const uint32_t* keys = reinterpret_cast<const uint32_t*>(pr.t_slots);
        int best_key = 0x7FFFFFFF;   // (d << 16 | i1): minimum distance, lowest query index = the last record setter
        for (int a = b; a < e; ++a) {
            const uint32_t ka = keys[a], i1 = ka >> 16, d = ka & 0x1FFu;
            bool seen = true;
            if (prm.best_lr) {
                for (int o = b; o < e; ++o) {
                    const uint32_t ko = keys[o];
                    if ((ko >> 16) < i1 && (ko & 0x1FFu) <= d) seen = false;
                }
            }
            pr.seen[ka >> 9] = seen ? 1 : 0;            // ka >> 9 = query * CAP + slot
            best_key = min(best_key, (int)((d << 16) | i1));
        }
        pr.m21[t] = (e > b) ? (best_key & 0xFFFF) : -1;
    }
    __syncthreads();

    // ---- D. best / second best, ratio test (:152-163), mutual filter (:166-174) ----
    int local = 0;
    for (int q = tid; q < n1; q += MG_THREADS) {
        const int2* seg = pr.q_pairs + (size_t)q * CAP;
        const int k = pr.q_count[q];
        int best_d = 0x7FFFFFFF, best_d2 = 0x7FFFFFFF, best_idx = -1;
        for (int j = 0; j < k; ++j) {
            const int d = seg[j].y;
            if (d < 0 || !pr.seen[(size_t)q * CAP + j]) continue;
            if (d < best_d) { best_d2 = best_d; best_d = d; best_idx = seg[j].x; }
            else if (d < best_d2) best_d2 = d;
        }
        int i2 = -1;
        if (k > 0 && (double)best_d < (double)best_d2 * prm.ratio) {   // int * double (:160)
            i2 = best_idx;
            local++;                                                   // matches++ (:162), whatever best_idx is
        }
        if (prm.best_lr && i2 >= 0 && pr.m21[i2] != q) {
            i2 = -1;
            local--;
        }
        pr.m12[q] = i2;
    }
    // count
    s_tmp[tid] = local;
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
        for (int i = 0; i < MG_THREADS; i++) tot += s_tmp[i];
        if (pr.count) *pr.count = tot;
    }
}

}  // namespace

size_t match_grid_smem_bytes(int rows, int cols) { return ((size_t)2 * rows * cols + 1 + MG_THREADS + 1) * sizeof(int); }

cudaError_t launch_match_grid(const GridProblem* problems, int B, const GridParams& prm, bool lines, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    if (prm.cap != MG_CAP) return cudaErrorInvalidValue;   // the scratch arrays are laid out for MG_CAP candidates per query
    const size_t smem = match_grid_smem_bytes(prm.rows, prm.cols);
    if (lines) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(match_grid_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        match_grid_kernel<true><<<B, MG_THREADS, smem, stream>>>(problems, prm);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(match_grid_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        match_grid_kernel<false><<<B, MG_THREADS, smem, stream>>>(problems, prm);
    }
    return cudaGetLastError();
}

}  // namespace plstvo
