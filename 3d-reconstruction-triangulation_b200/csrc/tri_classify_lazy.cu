// tri_classify_lazy.cu -- DroneClassifier::classifyDrones (src/DroneClassifier.cpp:96-332) WITHOUT enumerating the
// candidate combinations: the classifier for 17..32 cameras (and, on request, for any camera count).
//
// The reference's fillCombinationQueue (:156-198) pushes every complete node of a tree whose size grows like
// n_drones * 2^cameras: fine at 8 cameras (~1e3 leaves per frame), out of reach at 32 -- for the reference
// itself, for the oracle, and for the enumerating kernels of tri_classify.cu.  But classifyDrones never needs the queue,
// only, a handful of times per frame, "the first element the queue would pop that passes a test":
//   phase 1 (:241-246)  the best leaf made of detections inside a path's ray gate, unique against the used ones, with
//                       error < error_ and its point within MAX_STEP of the path's last point;
//   phase 2 (:204-213)  repeatedly the best leaf unique against everything used or kept so far.
// "Best" is Combination::operator< (:12-20): fewest unused cameras, then smallest error (ties: DFS order, as in
// tri_classify.cu).  A leaf's number of used cameras can only grow along a branch, so that element is found by a
// depth-first branch and bound over the SAME tree: children that take a detection first, a branch abandoned when even
// taking a detection on every remaining camera cannot reach the best count found so far, and -- exactly as the reference
// prunes (:185-187) -- below any node whose error exceeds error_.  Every node is solved with the arithmetic of the
// enumerating path (ref::dlt_point's accumulation order, ref::lm_point), so at camera counts both can run the two
// classifiers return bit-identical paths, assignments and phases (tests/test_gpu_classify.py at 8 and 12 cameras).
// The search is exact; what is bounded is its budget: a search that visits more than LAZY_NODE_BUDGET nodes
// latches TRI_ERR_CAPACITY instead of returning something else than the reference's answer.
//
// One CTA of LZ_WARPS warps per sequence.  Phase 1 is speculative as in tri_classify.cu: every tracked path's search runs on
// its own warp inside the path's own ray gate, ignoring the other paths' picks; warp 0 then confirms the picks in path order
// and searches a colliding path again with the used detections taken out.  Inside a search the control flow is warp-uniform
// and the lanes evaluate the children of a node in parallel (lane k <-> detection k of the node's camera), each from the
// parent's accumulated normal equations.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "tri_classify.cuh"

namespace tri {

constexpr int LZ_CAMS = TRI_MAX_CAMS;          // 32
constexpr int LZ_SLOTS = TRI_MAX_DETS + 1;     // children of a node: "none" + up to 15 detections
constexpr long long LAZY_NODE_BUDGET = 1 << 20;  // nodes one best-leaf search may visit

constexpr int LZ_WARPS = 6;  // paths searched at once in phase 1

// the frame: detections and their pixel rays (read by every warp)
struct LzFrame {
  int n[LZ_CAMS];
  double px[LZ_CAMS][TRI_MAX_DETS], py[LZ_CAMS][TRI_MAX_DETS];
  double dir[LZ_CAMS][TRI_MAX_DETS][3];
};
// one best-leaf search (one per warp)
struct LzSearch {
  unsigned short allowed[LZ_CAMS];  // detections the search may take
  int potential[LZ_CAMS + 1];       // cameras >= l with an allowed detection
  // state of the node at level l (cameras 0 .. l-1 decided)
  double M[LZ_CAMS + 1][6], v[LZ_CAMS + 1][3], X[LZ_CAMS + 1][3], err[LZ_CAMS + 1];
  int count[LZ_CAMS + 1];
  double row[LZ_CAMS][8];           // the two rows (a0 a1 a2 b) of the k-th selected detection, selection order
  int sel_cam[LZ_CAMS], sel_det[LZ_CAMS];
  unsigned char choice[LZ_CAMS];
  unsigned todo[LZ_CAMS];           // children of level l still to visit (bit 0 = none, bit k = detection k - 1)
  double c_err[LZ_CAMS][LZ_SLOTS], c_X[LZ_CAMS][LZ_SLOTS][3];
  // the best leaf so far
  int best_count;
  double best_err, best_X[3];
  unsigned char best_choice[LZ_CAMS];
};
// the linking state of the frame (warp 0, and the hand-over from the speculative searches)
struct LzLink {
  unsigned short all[LZ_CAMS];      // bit d: detection d exists
  unsigned short used[LZ_CAMS];     // detections of the combinations accepted so far in this frame
  unsigned short gate[TRI_MAX_DRONES][LZ_CAMS];
  // phase 1, speculative: each tracked path's best leaf inside its gate, the other paths' picks ignored
  int spec_count[TRI_MAX_DRONES];
  double spec_X[TRI_MAX_DRONES][3];
  unsigned char spec_choice[TRI_MAX_DRONES][LZ_CAMS];
  int abort;                        // a search ran out of its node budget
  unsigned processed;               // phase 1's confirmed paths
  int incumbent;                    // phase 2: the best count any warp of the shared search has found
  int more;                         // phase 2: another combination was kept
  // classifyPaths
  double fin_X[LINK_MAX_FINAL][3];
  unsigned char fin_choice[LINK_MAX_FINAL][LZ_CAMS];
  double pdist[LINK_MAX_FINAL][TRI_MAX_DRONES];
  int cp_path[LINK_MAX_FINAL];
  double cp_err[LINK_MAX_FINAL];
  LinkState S;
};
constexpr size_t LZ_SMEM_BYTES = sizeof(LzFrame) + sizeof(LzLink) + LZ_WARPS * sizeof(LzSearch);

// The best leaf (fewest unused cameras, then smallest error, then DFS order) whose detections lie in sh.allowed, whose error
// is < error_ and -- if `near` -- whose point is within MAX_STEP of near[0..2].  Result in sh.best_* (best_count = 0: none).
// Returns false if the node budget ran out.
__device__ bool lazy_best_leaf(const DltRig<double>& dlt, const RayRig& ray, const ClsParams& p, const LzFrame& fr, LzSearch& sh, const double* near,
                               unsigned long long& nodes, unsigned long long& solves, unsigned long long& lm_iters, int part = 0, int nparts = 1,
                               int* incumbent = nullptr) {
  const int lane = threadIdx.x & 31, C = p.n_cams;
  if (lane == 0) {
    int run = 0;
    sh.potential[C] = 0;
    for (int c = C - 1; c >= 0; c--) { run += sh.allowed[c] ? 1 : 0; sh.potential[c] = run; }
    sh.best_count = 0;
    sh.best_err = 0;
    sh.count[0] = 0;
    sh.err[0] = 0;
    for (int k = 0; k < 6; k++) sh.M[0][k] = 0;
    for (int k = 0; k < 3; k++) { sh.v[0][k] = 0; sh.X[0][k] = 0; }
  }
  __syncwarp();
  int l = 0;
  bool expanded = false;
  long long visited = 0;
  while (l >= 0) {
    if (!expanded) {
      // ---- expand the node at level l: lane k evaluates "camera l takes detection k - 1" from the parent's sums ----
      if (++visited > LAZY_NODE_BUDGET) return false;
      const int cnt = sh.count[l];
      unsigned ok_bit = 0;
      if (lane >= 1 && lane <= fr.n[l] && (sh.allowed[l] >> (lane - 1) & 1)) {
        const double x = fr.px[l][lane - 1], y = fr.py[l][lane - 1];
        double X[3] = {0, 0, 0}, e = 0;
        bool keep = true;
        if (p.solver == 0) {
          // ref::dlt_point's arithmetic, the sums of the first cnt detections taken from the parent (same order, same bits)
          const double* P = dlt.P[l];
          double M[6], v[3];
          for (int k = 0; k < 6; k++) M[k] = sh.M[l][k];
          for (int k = 0; k < 3; k++) v[k] = sh.v[l][k];
          double r[8];
          r[0] = P[0] - x * P[8]; r[1] = P[1] - x * P[9]; r[2] = P[2] - x * P[10]; r[3] = x * P[11] - P[3];
          r[4] = P[4] - y * P[8]; r[5] = P[5] - y * P[9]; r[6] = P[6] - y * P[10]; r[7] = y * P[11] - P[7];
          for (int h = 0; h < 2; h++) {
            const double a0 = r[4 * h], a1 = r[4 * h + 1], a2 = r[4 * h + 2], b = r[4 * h + 3];
            M[0] += a0 * a0; M[1] += a0 * a1; M[2] += a0 * a2; M[3] += a1 * a1; M[4] += a1 * a2; M[5] += a2 * a2;
            v[0] += a0 * b; v[1] += a1 * b; v[2] += a2 * b;
          }
          if (cnt + 1 >= 2) {
            solve_sym3<double>(M, v, X);
            double ss = 0;
            for (int i = 0; i < cnt; i++) {
              const double* q = sh.row[i];
              const double e0 = (q[0] * X[0] + q[1] * X[1] + q[2] * X[2]) - q[3];
              const double e1 = (q[4] * X[0] + q[5] * X[1] + q[6] * X[2]) - q[7];
              ss += e0 * e0;
              ss += e1 * e1;
            }
            const double e0 = (r[0] * X[0] + r[1] * X[1] + r[2] * X[2]) - r[3];
            const double e1 = (r[4] * X[0] + r[5] * X[1] + r[6] * X[2]) - r[7];
            ss += e0 * e0;
            ss += e1 * e1;
            e = sqrt(ss / (2 * (cnt + 1)));
            solves++;
            keep = !(e > p.error_);  // :185-187
          }
        } else if (cnt + 1 >= 2) {
          ref::RaySet rs;
          rs.n = cnt + 1;
          for (int i = 0; i < cnt; i++) {
            rs.cam[i] = sh.sel_cam[i];
            for (int j = 0; j < 3; j++) rs.d[i][j] = fr.dir[sh.sel_cam[i]][sh.sel_det[i]][j];
          }
          rs.cam[cnt] = l;
          for (int j = 0; j < 3; j++) rs.d[cnt][j] = fr.dir[l][lane - 1][j];
          int it = 0;
          e = solve_rays(ray, p.solver, rs, X, it);
          lm_iters += it;
          solves++;
          keep = !(e > p.error_);
        }
        sh.c_err[l][lane] = e;
        sh.c_X[l][lane][0] = X[0]; sh.c_X[l][lane][1] = X[1]; sh.c_X[l][lane][2] = X[2];
        ok_bit = keep ? 1u << lane : 0u;
      }
      unsigned children = 0;
      for (int o = 16; o > 0; o >>= 1) ok_bit |= __shfl_xor_sync(0xffffffffu, ok_bit, o);
      children = ok_bit | 1u;  // "none" repeats the parent's subset: same error, already known to pass
      if (l == 0 && nparts > 1) {
        // a search shared by several warps: this one takes every nparts-th subtree of the root (dealing out the subtrees of the
        // second level instead balances better but shares the bound later: measured slower, profiles/r2_link_kernel.log)
        unsigned mine = 0;
        int r = 0;
        for (unsigned m = children; m; r++) {
          const int k = 31 - __clz(m);
          m &= ~(1u << k);
          if (r % nparts == part) mine |= 1u << k;
        }
        children = mine;
      }
      if (lane == 0) sh.todo[l] = children;
      nodes += __popc(children);
      __syncwarp();
      expanded = true;
    }
    // ---- the next child of level l: detections first (they lead to the leaves with the fewest unused cameras), "none" last ----
    const unsigned todo = sh.todo[l];
    if (!todo) { l--; expanded = true; continue; }  // back to the parent, whose todo mask is still in shared memory
    // (the order of the visits does not change what the search returns -- ties are settled by the choices themselves -- but
    // visiting the child with the smallest error first finds a good leaf early, and the error bound below then cuts the rest)
    int k = 0;  // 0: none
    if (todo & ~1u) {
      const bool mine = lane >= 1 && lane < LZ_SLOTS && (todo >> lane & 1u);
      const u64 key = mine ? (u64)__double_as_longlong(sh.c_err[l][lane]) : ~0ull;  // errors are >= +0: their bits order like the values
      const unsigned hi = (unsigned)(key >> 32), mh = __reduce_min_sync(0xffffffffu, hi);
      const unsigned lo = hi == mh ? (unsigned)key : 0xffffffffu, ml = __reduce_min_sync(0xffffffffu, lo);
      k = 31 - __clz(__ballot_sync(0xffffffffu, mine && hi == mh && lo == ml));  // (equal errors: the highest detection, as before)
    }
    __syncwarp();
    if (lane == 0) sh.todo[l] = todo & ~(1u << k);
    const int cnt = sh.count[l] + (k ? 1 : 0);
    // bound: taking a detection on every remaining camera that has an allowed one
    // (the other warps' best count prunes too: a branch is only dropped when it cannot even tie, so what is found does not depend on timing)
    const int inc = incumbent ? *reinterpret_cast<volatile int*>(incumbent) : 0;
    if (cnt + sh.potential[l + 1] < max(max(sh.best_count, inc), MIN_CAMERAS)) { __syncwarp(); continue; }
    // second bound (matrix mode): a branch that can only tie the best count K must beat the best error.  The sum of squared
    // residuals at the least-squares point cannot shrink when rows are added, so every leaf below this child has
    // error^2 >= (child error)^2 cnt / K.  Only clear cases are dropped -- 1e-6 relative and 1e-10 absolute cover the rounding
    // of the sums, which are not computed at the exact minimiser --, so what the search returns is unchanged.
    if (p.solver == 0 && sh.best_count >= MIN_CAMERAS && inc <= sh.best_count && cnt + sh.potential[l + 1] == sh.best_count) {
      const double ec = k ? sh.c_err[l][k] : sh.err[l];
      if (ec * ec * (double)cnt > sh.best_err * sh.best_err * (double)sh.best_count * (1.0 + 1e-6) + 1e-10) { __syncwarp(); continue; }
    }
    if (lane == 0) {
      sh.choice[l] = (unsigned char)k;
      sh.count[l + 1] = cnt;
      if (k) {
        const int i = sh.count[l];
        sh.sel_cam[i] = l; sh.sel_det[i] = k - 1;
        sh.err[l + 1] = sh.c_err[l][k];
        for (int j = 0; j < 3; j++) sh.X[l + 1][j] = sh.c_X[l][k][j];
        if (p.solver == 0) {
          const double* P = dlt.P[l];
          const double x = fr.px[l][k - 1], y = fr.py[l][k - 1];
          double* r = sh.row[i];
          r[0] = P[0] - x * P[8]; r[1] = P[1] - x * P[9]; r[2] = P[2] - x * P[10]; r[3] = x * P[11] - P[3];
          r[4] = P[4] - y * P[8]; r[5] = P[5] - y * P[9]; r[6] = P[6] - y * P[10]; r[7] = y * P[11] - P[7];
          double M[6], v[3];
          for (int q = 0; q < 6; q++) M[q] = sh.M[l][q];
          for (int q = 0; q < 3; q++) v[q] = sh.v[l][q];
          for (int h = 0; h < 2; h++) {
            const double a0 = r[4 * h], a1 = r[4 * h + 1], a2 = r[4 * h + 2], b = r[4 * h + 3];
            M[0] += a0 * a0; M[1] += a0 * a1; M[2] += a0 * a2; M[3] += a1 * a1; M[4] += a1 * a2; M[5] += a2 * a2;
            v[0] += a0 * b; v[1] += a1 * b; v[2] += a2 * b;
          }
          for (int q = 0; q < 6; q++) sh.M[l + 1][q] = M[q];
          for (int q = 0; q < 3; q++) sh.v[l + 1][q] = v[q];
        }
      } else {
        sh.err[l + 1] = sh.err[l];
        for (int j = 0; j < 3; j++) sh.X[l + 1][j] = sh.X[l][j];
        for (int q = 0; q < 6; q++) sh.M[l + 1][q] = sh.M[l][q];
        for (int q = 0; q < 3; q++) sh.v[l + 1][q] = sh.v[l][q];
      }
    }
    __syncwarp();
    if (l + 1 < C) { l++; expanded = false; continue; }
    // ---- a leaf (:190-196): a candidate if it has >= MIN_CAMERAS detections, error < error_ and passes the caller's test ----
    if (cnt >= MIN_CAMERAS && sh.err[C] < p.error_) {
      bool pass = true;
      if (near) {  // cv::norm(c.point - pos) < MAX_STEP, :244
        const double dx = sh.X[C][0] - near[0], dy = sh.X[C][1] - near[1], dz = sh.X[C][2] - near[2];
        pass = sqrt(dx * dx + dy * dy + dz * dz) < MAX_STEP;
      }
      if (pass) {
        bool better = cnt > sh.best_count || (cnt == sh.best_count && sh.err[C] < sh.best_err);
        if (!better && cnt == sh.best_count && sh.err[C] == sh.best_err) {  // a tie: DFS order = lexicographic in the choices, "none" first
          int first = C;
          for (int c0 = 0; c0 < C; c0 += 32) {
            const int c = c0 + lane;
            const unsigned diff = __ballot_sync(0xffffffffu, c < C && sh.choice[c] != sh.best_choice[c]);
            if (diff) { first = c0 + __ffs(diff) - 1; break; }
          }
          better = first < C && sh.choice[first] < sh.best_choice[first];
        }
        if (better) {
          __syncwarp();
          if (lane == 0) { sh.best_count = cnt; sh.best_err = sh.err[C]; for (int j = 0; j < 3; j++) sh.best_X[j] = sh.X[C][j]; if (incumbent) atomicMax(incumbent, cnt); }
          for (int c = lane; c < C; c += 32) sh.best_choice[c] = sh.choice[c];
        }
      }
    }
    __syncwarp();
    expanded = true;  // stay on level l: its todo mask decides
  }
  return true;
}

__global__ void __launch_bounds__(32 * LZ_WARPS)
lazy_link_kernel(const __grid_constant__ DltRig<double> dlt, const __grid_constant__ RayRig ray, ClsParams p, const int2* __restrict__ seq_bounds,
                 const int32_t* __restrict__ offs, const double* __restrict__ dets, LinkState* state, double* __restrict__ out_paths,
                 int8_t* __restrict__ out_assign, uint8_t* __restrict__ out_phase, ClsCounters* ctr) {
  extern __shared__ __align__(16) unsigned char lazy_dyn[];
  LzFrame& fr = *reinterpret_cast<LzFrame*>(lazy_dyn);
  LzLink& lk = *reinterpret_cast<LzLink*>(lazy_dyn + sizeof(LzFrame));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, C = p.n_cams, D = p.n_drones;
  LzSearch& sh = reinterpret_cast<LzSearch*>(lazy_dyn + sizeof(LzFrame) + sizeof(LzLink))[warp];
  const int fa = seq_bounds ? seq_bounds[blockIdx.x].x : p.f0, fb = seq_bounds ? seq_bounds[blockIdx.x].y : p.f1;
  LinkState* st = state + blockIdx.x;
  for (int i = tid; i < (int)(sizeof(LinkState) / sizeof(int)); i += 32 * LZ_WARPS) ((int*)&lk.S)[i] = ((const int*)st)[i];
  if (tid == 0) lk.abort = 0;
  __syncthreads();
  unsigned long long nodes = 0, solves = 0, lm_iters = 0, n_phase1 = 0, n_phase2 = 0, leaves = 0;
  bool overflow_final = false, bad_input = false;

  auto emit = [&](int path, int f, const double* X, const unsigned char* choice, int phase) {  // warp 0, lane 0
    const int n = lk.S.n[path];
    double(*t)[3] = lk.S.tail[path];
    if (n >= PATH_TAIL) {
      for (int k = 0; k < PATH_TAIL - 1; k++) for (int j = 0; j < 3; j++) t[k][j] = t[k + 1][j];
      for (int j = 0; j < 3; j++) t[PATH_TAIL - 1][j] = X[j];
    } else {
      for (int j = 0; j < 3; j++) t[n][j] = X[j];
    }
    if (n < 0x3fffffff) lk.S.n[path] = n + 1;
    double* o = out_paths + ((size_t)path * p.n_frames + f) * 3;
    o[0] = X[0]; o[1] = X[1]; o[2] = X[2];
    if (out_assign) for (int c = 0; c < C; c++) out_assign[((size_t)path * p.n_frames + f) * C + c] = (int8_t)choice[c];
    if (out_phase) out_phase[(size_t)path * p.n_frames + f] = (uint8_t)phase;
  };
  // the MAX_STEP ray gate of a path (:228-236), lane <-> camera, into lk.gate[np]; allowed = gate & ~(used if filtered); returns
  // the number of cameras with an allowed detection
  auto gate_path = [&](int np, const double* last, bool filtered, bool compute) {
    int cams = 0;
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int c = c0 + lane;
      unsigned g = 0;
      if (c < C) {
        if (compute) {
          for (int d = 0; d < fr.n[c]; d++)
            if (ref::dist_to_ray(ray.pos[c], fr.dir[c][d], last[0], last[1], last[2]) < MAX_STEP) g |= 1u << d;
          lk.gate[np][c] = (unsigned short)g;
        } else {
          g = lk.gate[np][c];
        }
        if (filtered) g &= ~(unsigned)lk.used[c];
        sh.allowed[c] = (unsigned short)g;
      }
      cams += __popc(__ballot_sync(0xffffffffu, c < C && g != 0));
    }
    __syncwarp();
    return cams;
  };

  for (int f = fa; f < fb; f++) {
    // ---- the frame's detections and their pixel rays (all warps) ----
    for (int c = tid; c < C; c += 32 * LZ_WARPS) {
      const int a = offs[(size_t)c * (p.n_frames + 1) + f], b = offs[(size_t)c * (p.n_frames + 1) + f + 1];
      if (b < a || b - a > TRI_MAX_DETS) bad_input = true;
      const int n = min(max(b - a, 0), TRI_MAX_DETS);
      fr.n[c] = n;
      lk.all[c] = (unsigned short)((1u << n) - 1u);
      lk.used[c] = 0;
    }
    __syncthreads();
    for (int i = tid; i < C * TRI_MAX_DETS; i += 32 * LZ_WARPS) {
      const int c = i / TRI_MAX_DETS, d = i - c * TRI_MAX_DETS;
      if (d < fr.n[c]) {
        const int a = offs[(size_t)c * (p.n_frames + 1) + f];
        const double x = dets[2 * (size_t)(a + d)], y = dets[2 * (size_t)(a + d) + 1];
        fr.px[c][d] = x; fr.py[c][d] = y;
        ref::make_dir(ray, c, x, y, fr.dir[c][d]);
      }
    }
    __syncthreads();

    // ---- phase 1, speculative: every tracked path's best leaf inside its own gate, the other paths' picks ignored;
    // LZ_WARPS paths at a time, one warp each (:121-123 decides which paths track) ----
    unsigned act_mask = 0;
    for (int np = 0; np < D; np++) {
      const int n = lk.S.n[np];
      if (n == 0) continue;
      const double* last = lk.S.tail[np][min(n, PATH_TAIL) - 1];
      if (last[0] == 0 && last[1] == 0 && last[2] == 0) continue;
      act_mask |= 1u << np;
    }
    {
      int ai = 0;
      for (unsigned rem = act_mask; rem; rem &= rem - 1, ai++) {
        if (ai % LZ_WARPS != warp) continue;
        const int np = __ffs(rem) - 1;
        const double* last = lk.S.tail[np][min(lk.S.n[np], PATH_TAIL) - 1];
        const int cams_in_gate = gate_path(np, last, false, true);
        int found = 0;
        if (cams_in_gate >= MIN_CAMERAS) {
          if (!lazy_best_leaf(dlt, ray, p, fr, sh, last, nodes, solves, lm_iters)) { if (lane == 0) lk.abort = 1; }
          else found = sh.best_count;
        }
        if (lane == 0) { lk.spec_count[np] = found; for (int j = 0; j < 3; j++) lk.spec_X[np][j] = sh.best_X[j]; }
        if (found >= MIN_CAMERAS) for (int c = lane; c < C; c += 32) lk.spec_choice[np][c] = sh.best_choice[c];
      }
    }
    __syncthreads();
    if (lk.abort) break;

    if (warp == 0) {
      // ---- confirm in path order (:119-135): a pick that shares no detection with the ones accepted before it is also the
      // best leaf of the filtered search; a colliding one is searched again with the used detections taken out ----
      unsigned processed = 0;
      for (unsigned rem = act_mask; rem; rem &= rem - 1) {
        const int np = __ffs(rem) - 1;
        if (lk.spec_count[np] < MIN_CAMERAS) continue;  // nothing inside the gate even with every detection free
        bool clash = false;
        for (int c0 = 0; c0 < C; c0 += 32) {
          const int c = c0 + lane;
          const int k = c < C ? lk.spec_choice[np][c] : 0;
          clash = clash || __any_sync(0xffffffffu, k && (lk.used[c < C ? c : 0] >> (k - 1) & 1));
        }
        const double* X = lk.spec_X[np];
        const unsigned char* choice = lk.spec_choice[np];
        if (clash) {
          const double* last = lk.S.tail[np][min(lk.S.n[np], PATH_TAIL) - 1];
          if (gate_path(np, last, true, false) < MIN_CAMERAS) continue;
          if (!lazy_best_leaf(dlt, ray, p, fr, sh, last, nodes, solves, lm_iters)) { if (lane == 0) lk.abort = 1; break; }
          if (sh.best_count < MIN_CAMERAS) continue;
          X = sh.best_X; choice = sh.best_choice;
        }
        leaves++;
        processed |= 1u << np;
        for (int c = lane; c < C; c += 32)
          if (choice[c]) lk.used[c] |= (unsigned short)(1u << (choice[c] - 1));
        __syncwarp();
        if (lane == 0) { emit(np, f, X, choice, 1); n_phase1++; }
        __syncwarp();
      }
      if (lane == 0) lk.processed = processed;
    }
    __syncthreads();
    if (lk.abort) break;
    const unsigned processed = lk.processed;

    if (__popc(processed) != D) {  // :137
      // ---- phase 2: pickBestCombinations (:200-217): the best leaf among the unused detections, again and again.  One search,
      // all warps: each takes every LZ_WARPS-th subtree of the root, the best count found so far is shared ----
      int n_fin = 0;
      for (;;) {
        int live = 0;
        for (int c0 = 0; c0 < C; c0 += 32) {
          const int c = c0 + lane;
          if (c < C) sh.allowed[c] = (unsigned short)(lk.all[c] & ~lk.used[c]);
          live += __popc(__ballot_sync(0xffffffffu, c < C && (lk.all[c] & ~lk.used[c]) != 0));
        }
        if (tid == 0) { lk.incumbent = 0; if (n_fin == 0) lk.more = 0; }
        __syncthreads();
        if (live < MIN_CAMERAS) break;
        if (!lazy_best_leaf(dlt, ray, p, fr, sh, nullptr, nodes, solves, lm_iters, warp, LZ_WARPS, &lk.incumbent) && lane == 0) lk.abort = 1;
        __syncthreads();
        if (lk.abort) break;
        if (warp == 0) {  // the best of the warps' bests: most cameras, smallest error, then DFS order (lexicographic in the choices)
          LzSearch* all = reinterpret_cast<LzSearch*>(lazy_dyn + sizeof(LzFrame) + sizeof(LzLink));
          int wb = -1;
          for (int w = 0; w < LZ_WARPS; w++) {
            const LzSearch& q = all[w];
            if (q.best_count < MIN_CAMERAS) continue;
            bool better = wb < 0 || q.best_count > all[wb].best_count || (q.best_count == all[wb].best_count && q.best_err < all[wb].best_err);
            if (!better && q.best_count == all[wb].best_count && q.best_err == all[wb].best_err) {
              int first = C;
              for (int c0 = 0; c0 < C; c0 += 32) {
                const int c = c0 + lane;
                const unsigned diff = __ballot_sync(0xffffffffu, c < C && q.best_choice[c] != all[wb].best_choice[c]);
                if (diff) { first = c0 + __ffs(diff) - 1; break; }
              }
              better = first < C && q.best_choice[first] < all[wb].best_choice[first];
            }
            if (better) wb = w;
          }
          if (wb >= 0) {
            leaves++;
            if (n_fin >= LINK_MAX_FINAL) overflow_final = true;
            else {
              const LzSearch& q = all[wb];
              if (lane == 0) { for (int j = 0; j < 3; j++) lk.fin_X[n_fin][j] = q.best_X[j]; lk.more = n_fin + 1; }  // (only ever rises within a frame)
              for (int c = lane; c < C; c += 32) {
                lk.fin_choice[n_fin][c] = q.best_choice[c];
                if (q.best_choice[c]) lk.used[c] |= (unsigned short)(1u << (q.best_choice[c] - 1));
              }
            }
          }
        }
        __syncthreads();
        if (lk.more != n_fin + 1) break;
        n_fin++;
      }
      if (!lk.abort && warp == 0) {
        // ---- classifyPaths (:262-332) ----
        unsigned open_paths = 0;
        for (int j = 0; j < D; j++) if (!(processed >> j & 1u) && lk.S.n[j] != 0) open_paths |= 1u << j;
        const int n_open = __popc(open_paths);
        for (int q = lane; q < n_fin * n_open; q += 32) {
          const int ci = q / n_open;
          int j = 0;
          { unsigned rem = open_paths; for (int s2 = q - ci * n_open; s2 > 0; s2--) rem &= rem - 1; j = __ffs(rem) - 1; }
          const int npc = min(lk.S.n[j], PATH_TAIL);
          double dist = 0;
          for (int t = 0; t < npc; t++) {
            const double dx = lk.S.tail[j][t][0] - lk.fin_X[ci][0], dy = lk.S.tail[j][t][1] - lk.fin_X[ci][1], dz = lk.S.tail[j][t][2] - lk.fin_X[ci][2];
            dist += sqrt(dx * dx + dy * dy + dz * dz);
          }
          lk.pdist[ci][j] = dist / (double)npc;
        }
        __syncwarp();
        for (int i = lane; i < n_fin; i += 32) {
          int bestPath = 0;
          double bestDist = -1;
          for (unsigned rem = open_paths; rem; rem &= rem - 1) {
            const int j = __ffs(rem) - 1;
            const double dist = lk.pdist[i][j];
            if (dist < bestDist || bestDist == -1) { bestDist = dist; bestPath = j; }
          }
          lk.cp_path[i] = bestPath; lk.cp_err[i] = bestDist;
        }
        __syncwarp();
        if (lane == 0) {  // the stable ascending order of :299-300, walked as :302-321 does
          unsigned done = processed;
          unsigned long long taken_lo = 0, taken_hi = 0;
          for (int step = 0; step < n_fin; step++) {
            int idx = -1;
            for (int i = 0; i < n_fin; i++) {
              if ((i < 64 ? taken_lo >> i : taken_hi >> (i - 64)) & 1ull) continue;
              if (idx < 0 || lk.cp_err[i] < lk.cp_err[idx]) idx = i;
            }
            if (idx < 64) taken_lo |= 1ull << idx; else taken_hi |= 1ull << (idx - 64);
            int target = -1;
            if (done >> lk.cp_path[idx] & 1u) {
              for (int i = 0; i < D; i++) if (lk.S.n[i] == 0) { target = i; break; }
            } else {
              target = lk.cp_path[idx];
            }
            if (target != -1) { emit(target, f, lk.fin_X[idx], lk.fin_choice[idx], 2); done |= 1u << target; n_phase2++; }
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();  // the frame is linked: its data may be overwritten, the state is current
    if (lk.abort) break;
  }
  __syncthreads();
  for (int i = tid; i < (int)(sizeof(LinkState) / sizeof(int)); i += 32 * LZ_WARPS) ((int*)st)[i] = ((const int*)&lk.S)[i];
  for (int o = 16; o > 0; o >>= 1) {
    solves += __shfl_down_sync(0xffffffffu, solves, o); lm_iters += __shfl_down_sync(0xffffffffu, lm_iters, o);
  }
  if (lane == 0) {
    atomicAdd(&ctr->nodes, nodes); atomicAdd(&ctr->solves, solves); atomicAdd(&ctr->lm_iters, lm_iters);
    atomicAdd(&ctr->leaves, leaves); atomicAdd(&ctr->phase1, n_phase1); atomicAdd(&ctr->phase2, n_phase2);
    if (overflow_final) atomicExch(&ctr->overflow_final, 1);
  }
  if (tid == 0 && lk.abort) atomicExch(&ctr->overflow_frontier, 1);
  if (__any_sync(0xffffffffu, bad_input) && lane == 0) atomicExch(&ctr->bad_input, 1);
}

cudaError_t launch_lazy_link(cudaStream_t s, const DltRig<double>& dlt, const RayRig& ray, const ClsParams& p, int n_seq, const int2* d_seq,
                             const int32_t* d_offs, const double* d_dets, LinkState* d_state, double* d_paths, int8_t* d_assign,
                             uint8_t* d_phase, ClsCounters* d_ctr) {
  cudaError_t err = cudaFuncSetAttribute(lazy_link_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LZ_SMEM_BYTES);
  if (err != cudaSuccess) return err;
  lazy_link_kernel<<<n_seq, 32 * LZ_WARPS, LZ_SMEM_BYTES, s>>>(dlt, ray, p, d_seq, d_offs, d_dets, d_state, d_paths, d_assign, d_phase, d_ctr);
  return cudaGetLastError();
}

}  // namespace tri
