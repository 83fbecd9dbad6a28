// tri_classify.cu -- kernel 3: DroneClassifier::classifyDrones (src/DroneClassifier.cpp:96-332) as two
// kernels.  Compiled with -fmad=false: every compare that decides an index (error_ pruning, the
// MAX_STEP gates, the priority order) is evaluated in the reference's operation order.
//
// (A) enumerate_kernel -- frame-parallel.  fillCombinationQueue (:156-198) walks a tree whose level c
//     picks, for camera c, "none" (0) or detection k (1..n_c); every node with >= 2 detections is
//     triangulated and its subtree cut when error > error_ (:185-187); complete nodes with
//     >= MIN_CAMERAS detections are the candidate combinations (:190-196).  Whether a node survives
//     depends only on its own prefix, so the tree is expanded level-synchronously: one CTA owns a
//     frame, a level's children are evaluated one per thread and compacted IN ORDER (ballot + scan),
//     which reproduces the reference's DFS order of the leaves (lexicographic in the per-camera
//     choices).  A combination is one 64-bit word, 4 bits per camera.  The gated enumerations of
//     triangulateWithLastPos (:219-250) are subsets of this full-frame leaf list (same prefixes, same
//     pixels => same errors), so the tree is expanded ONCE per frame instead of once per path.
// (B) link_kernel -- frame-sequential (the reference's tracking state: last position + 3-point tail
//     per path, :119-135, :269-297).  One CTA walks the frames in order; per path it evaluates the
//     MAX_STEP ray gate of every detection in parallel, then takes the arg-min over the admissible
//     leaves of the key (fewest unused cameras, smallest error, DFS order) -- the first element the
//     reference's priority_queue would pop that passes :241-246.  Phase 2 (pickBestCombinations +
//     classifyPaths, :200-217, :262-332) is the same arg-min repeated greedily, then the tiny
//     path-assignment logic on one thread.  Ties on (unused cameras, error) are broken by DFS order
//     (the reference's heap order is an artefact of libstdc++); they are counted in stats.ties.
#include <float.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "tri_engine.cuh"
#include "tri_ref.cuh"

namespace tri {

constexpr int CLS_MAX_CAMS = 16;   // 4 bits per camera in a 64-bit combination
constexpr int CLS_THREADS = 128;
constexpr int LINK_THREADS = 256;
constexpr int LINK_MAX_FINAL = 64;
constexpr int LINK_STAGE_LEAVES = 2048;  // leaves of one frame staged in (dynamic) shared memory, 40 B each; two buffers
typedef unsigned long long u64;

// DroneClassifier.h:11-17
constexpr double MAX_ERROR_MATRIX = 1e+5, MAX_ERROR_RAY = 120, MAX_STEP = 200;
constexpr int MIN_CAMERAS = 2, PATH_TAIL = 3;

struct ClsParams {
  int n_cams, n_drones, solver;  // solver: 0 matrix, 1 ray reference LM, 2 ray closed form
  int n_frames;                  // whole sequence (row length of the CSR offsets is n_frames + 1)
  int f0, f1;                    // frame batch [f0, f1)
  int cap;                       // frontier capacity per CTA
  long long leaf_cap;
  double error_;
};

struct ClsCounters {
  u64 leaf_total, nodes, solves, leaves, lm_iters, phase1, phase2, ties;
  int max_frontier, overflow_frontier, overflow_leaves, bad_input;
};

struct LinkState {
  double tail[TRI_MAX_DRONES][PATH_TAIL][3];  // oldest .. newest of the last min(n,3) points
  int n[TRI_MAX_DRONES];                      // points pushed so far (saturating)
};

__device__ __forceinline__ u64 nonzero_nibbles(u64 v) { return (v | (v >> 1) | (v >> 2) | (v >> 3)) & 0x1111111111111111ull; }

// isCombinationUnique (DroneClassifier.cpp:32-41): no camera where both use the same detection
__device__ __forceinline__ bool conflicts(u64 a, u64 b) { return (nonzero_nibbles(a) & ~nonzero_nibbles(a ^ b)) != 0; }

__device__ inline double solve_combination(const DltRig<double>& dlt, const RayRig& ray, int solver, u64 comb, int n_cams,
                                           const double (*px)[TRI_MAX_DETS], const double (*py)[TRI_MAX_DETS], double X[3],
                                           int& iters) {
  iters = 0;
  if (solver == 0) {
    int cam[CLS_MAX_CAMS], n = 0;
    double x[CLS_MAX_CAMS], y[CLS_MAX_CAMS];
    for (int i = 0; i < n_cams; i++) {
      const int k = (int)((comb >> (4 * i)) & 15);
      if (k) { cam[n] = i; x[n] = px[i][k - 1]; y[n] = py[i][k - 1]; n++; }
    }
    return ref::dlt_point(dlt, n, cam, x, y, X);
  }
  ref::RaySet rs;
  rs.n = 0;
  for (int i = 0; i < n_cams; i++) {
    const int k = (int)((comb >> (4 * i)) & 15);
    if (k) { rs.cam[rs.n] = i; ref::make_dir(ray, i, px[i][k - 1], py[i][k - 1], rs.d[rs.n]); rs.n++; }
  }
  if (solver == 1) return ref::lm_point(ray, rs, X, iters);
  // closed form about the mean origin (the minimiser the reference's LM converges to)
  const int n = rs.n;
  double m[3] = {0, 0, 0};
  for (int k = 0; k < n; k++) for (int j = 0; j < 3; j++) m[j] += ray.pos[rs.cam[k]][j] / n;
  double M[6] = {0, 0, 0, 0, 0, 0}, c[3] = {0, 0, 0};
  for (int k = 0; k < n; k++) {
    const double* d = rs.d[k];
    const double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    const double o[3] = {ray.pos[rs.cam[k]][0] - m[0], ray.pos[rs.cam[k]][1] - m[1], ray.pos[rs.cam[k]][2] - m[2]};
    const double dot = d[0] * o[0] + d[1] * o[1] + d[2] * o[2];
    M[0] += dd - d[0] * d[0]; M[1] -= d[0] * d[1]; M[2] -= d[0] * d[2];
    M[3] += dd - d[1] * d[1]; M[4] -= d[1] * d[2]; M[5] += dd - d[2] * d[2];
    for (int j = 0; j < 3; j++) c[j] += dd * o[j] - d[j] * dot;
  }
  solve_sym3<double>(M, c, X);
  for (int j = 0; j < 3; j++) X[j] += m[j];
  double S, rmax, e;
  ref::residual_pass(ray, rs, X, S, rmax, e);
  iters = 1;
  return e;
}

// ---- (A) candidate generation ------------------------------------------------------------------
__global__ void __launch_bounds__(CLS_THREADS)
enumerate_kernel(const __grid_constant__ DltRig<double> dlt, const __grid_constant__ RayRig ray, ClsParams p,
                 const int32_t* __restrict__ offs, const double* __restrict__ dets, u64* __restrict__ front,
                 double* __restrict__ tmp_xyz, double* __restrict__ tmp_err, u64* __restrict__ leaf_comb,
                 double* __restrict__ leaf_err, double* __restrict__ leaf_xyz, long long* __restrict__ leaf_off,
                 int* __restrict__ leaf_cnt, ClsCounters* ctr) {
  __shared__ int s_n[CLS_MAX_CAMS];
  __shared__ double s_px[CLS_MAX_CAMS][TRI_MAX_DETS], s_py[CLS_MAX_CAMS][TRI_MAX_DETS];
  __shared__ int s_warp[CLS_THREADS / 32];
  __shared__ long long s_off;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.n_cams;
  u64* buf0 = front + (size_t)blockIdx.x * 2 * p.cap;
  u64* buf1 = buf0 + p.cap;
  double* t_xyz = tmp_xyz + (size_t)blockIdx.x * 3 * p.cap;
  double* t_err = tmp_err + (size_t)blockIdx.x * p.cap;
  u64 st_nodes = 0, st_solves = 0, st_iters = 0;

  for (int f = p.f0 + blockIdx.x; f < p.f1; f += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < C * TRI_MAX_DETS; i += CLS_THREADS) {
      const int c = i / TRI_MAX_DETS, d = i % TRI_MAX_DETS;
      const int a = offs[(size_t)c * (p.n_frames + 1) + f], b = offs[(size_t)c * (p.n_frames + 1) + f + 1];
      if (d == 0) {
        s_n[c] = min(b - a, TRI_MAX_DETS);
        if (b - a > TRI_MAX_DETS || b < a) atomicExch(&ctr->bad_input, 1);
      }
      if (d < b - a) { s_px[c][d] = dets[2 * (size_t)(a + d)]; s_py[c][d] = dets[2 * (size_t)(a + d) + 1]; }
    }
    if (tid == 0) buf0[0] = 0;
    __syncthreads();
    u64 *fin = buf0, *fout = buf1;
    int m = 1;
    for (int c = 0; c < C && m > 0; c++) {
      const int nch = s_n[c] + 1;
      const int total = m * nch;
      const bool last = c == C - 1;
      int out_base = 0;
      for (int j0 = 0; j0 < total; j0 += CLS_THREADS) {
        const int j = j0 + tid;
        bool keep = false;
        u64 comb = 0;
        double X[3] = {0, 0, 0}, err = 0;
        if (j < total) {
          const int parent = j / nch, k = j - parent * nch;
          comb = fin[parent] | ((u64)k << (4 * c));
          const int cnt = __popcll(nonzero_nibbles(comb));
          keep = true;
          st_nodes++;
          if (cnt >= 2) st_solves++;  // the reference re-solves the "none" children too (:166-181)
          if (cnt >= 2 && (k > 0 || last)) {  // a "none" child repeats its parent's subset: same error
            int it;
            err = solve_combination(dlt, ray, p.solver, comb, c + 1, s_px, s_py, X, it);
            st_iters += it;
            if (err > p.error_) keep = false;  // :185-187
          }
          if (last && cnt < MIN_CAMERAS) keep = false;  // complete but too few detections: no candidate
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(ballot);
        __syncthreads();
        int before = 0, chunk = 0;
#pragma unroll
        for (int w = 0; w < CLS_THREADS / 32; w++) {
          if (w < warp) before += s_warp[w];
          chunk += s_warp[w];
        }
        const int pos = out_base + before + __popc(ballot & ((1u << lane) - 1));
        if (keep) {
          if (pos < p.cap) {
            fout[pos] = comb;
            if (last) { t_xyz[3 * pos] = X[0]; t_xyz[3 * pos + 1] = X[1]; t_xyz[3 * pos + 2] = X[2]; t_err[pos] = err; }
          } else {
            atomicExch(&ctr->overflow_frontier, 1);
          }
        }
        out_base += chunk;
        __syncthreads();
      }
      m = min(out_base, p.cap);
      if (tid == 0) atomicMax(&ctr->max_frontier, out_base);
      if (last && m > 64 * LINK_THREADS && tid == 0) atomicExch(&ctr->bad_input, 2);  // the linking pass keeps one live bit per leaf in a u64 per thread
      u64* t = fin; fin = fout; fout = t;
      if (m == 0) break;
      if (last) {
        // publish this frame's leaves (in order) into a contiguous range of the global leaf arrays
        if (tid == 0) {
          long long off = (long long)atomicAdd(&ctr->leaf_total, (u64)m);
          if (off + m > p.leaf_cap) { atomicExch(&ctr->overflow_leaves, 1); off = -1; }
          s_off = off;
        }
        __syncthreads();
        const long long off = s_off;
        if (off >= 0) {
          // ... in PRIORITY order: rank of leaf i = number of leaves the reference's priority_queue pops
          // before it = those with fewer unused cameras, then smaller error (Combination::operator<,
          // :12-20), ties by DFS order.  Rank by counting (m is ~1e3; frames are independent, so this is
          // parallel work) and scatter; the sequential linking kernel then only walks prefixes.
          const u64 cam_bits = C == 16 ? ~0ull : ((1ull << (4 * C)) - 1);
          u64 n_tie = 0;
          for (int i = tid; i < m; i += CLS_THREADS) {
            const u64 ci = fin[i];
            const double ei = t_err[i];
            const int zi = C - __popcll(nonzero_nibbles(ci & cam_bits));
            int rank = 0;
            bool tie = false;
            for (int j = 0; j < m; j++) {
              const int zj = C - __popcll(nonzero_nibbles(fin[j] & cam_bits));
              const double ej = t_err[j];
              const bool eq = zj == zi && ej == ei;
              rank += (zj < zi) || (zj == zi && ej < ei) || (eq && j < i);
              tie = tie || (eq && j != i);
            }
            n_tie += tie;
            leaf_comb[off + rank] = ci;
            leaf_err[off + rank] = ei;
            leaf_xyz[3 * (off + rank)] = t_xyz[3 * i]; leaf_xyz[3 * (off + rank) + 1] = t_xyz[3 * i + 1]; leaf_xyz[3 * (off + rank) + 2] = t_xyz[3 * i + 2];
          }
          if (n_tie) atomicAdd(&ctr->ties, n_tie);
          if (tid == 0) { leaf_off[f - p.f0] = off; leaf_cnt[f - p.f0] = m; atomicAdd(&ctr->leaves, (u64)m); }
        } else if (tid == 0) {
          leaf_off[f - p.f0] = 0; leaf_cnt[f - p.f0] = 0;
        }
        m = -1;  // done
      }
    }
    if (m >= 0 && tid == 0) { leaf_off[f - p.f0] = 0; leaf_cnt[f - p.f0] = 0; }  // the tree died out (or no cameras)
  }
  // block totals of the per-thread statistics
  for (int o = 16; o > 0; o >>= 1) {
    st_nodes += __shfl_down_sync(0xffffffffu, st_nodes, o);
    st_solves += __shfl_down_sync(0xffffffffu, st_solves, o);
    st_iters += __shfl_down_sync(0xffffffffu, st_iters, o);
  }
  if (lane == 0) { atomicAdd(&ctr->nodes, st_nodes); atomicAdd(&ctr->solves, st_solves); atomicAdd(&ctr->lm_iters, st_iters); }
}

// ---- (B) linking -------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ double dist3(const double* a, const double* b) {  // cv::norm(a - b)
  const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
  return sqrt(x * x + y * y + z * z);
}

__global__ void __launch_bounds__(LINK_THREADS)
link_kernel(const __grid_constant__ RayRig ray, ClsParams p, const int32_t* __restrict__ offs, const double* __restrict__ dets,
            const u64* __restrict__ leaf_comb, const double* __restrict__ leaf_err, const double* __restrict__ leaf_xyz,
            const long long* __restrict__ leaf_off, const int* __restrict__ leaf_cnt, LinkState* state,
            double* __restrict__ out_paths, int8_t* __restrict__ out_assign, uint8_t* __restrict__ out_phase, ClsCounters* ctr) {
  __shared__ LinkState S;
  __shared__ unsigned s_gate[CLS_MAX_CAMS][16];  // [camera][choice] -> bit np: path np may use that choice (choice 0 = "none")
  __shared__ unsigned s_active_mask, s_found_mask;
  __shared__ double s_dir[CLS_MAX_CAMS * TRI_MAX_DETS][3];
  __shared__ bool s_ndet[CLS_MAX_CAMS * TRI_MAX_DETS];
  __shared__ int s_cand[TRI_MAX_DRONES];
  __shared__ bool s_active[TRI_MAX_DRONES];
  __shared__ double s_pdist[LINK_MAX_FINAL][TRI_MAX_DRONES];
  __shared__ u64 s_used[TRI_MAX_DRONES];
  __shared__ u64 s_fin[LINK_MAX_FINAL];
  __shared__ int s_fin_idx[LINK_MAX_FINAL];
  __shared__ int s_n_fin, s_n_used;
  __shared__ int s_first[2][LINK_THREADS / 32];
  __shared__ unsigned s_processed;
  // This frame's candidates (combination, error, point) are staged in shared memory, and the NEXT frame's are
  // already on their way (cp.async into the other buffer) while this one is processed: a lone CTA cannot hide a
  // ~1 us global round trip behind anything else.
  extern __shared__ __align__(16) unsigned char link_dyn[];
  constexpr int LINK_BUF_BYTES = (int)((sizeof(u64) + 4 * sizeof(double)) * LINK_STAGE_LEAVES);
  const int tid = threadIdx.x, C = p.n_cams, D = p.n_drones;
  for (int i = tid; i < (int)(sizeof(LinkState) / sizeof(int)); i += LINK_THREADS) ((int*)&S)[i] = ((const int*)state)[i];
  u64 n_phase1 = 0, n_phase2 = 0;  // thread 0 only
  __syncthreads();

  auto emit = [&](int path, int f, u64 comb, const double* pt, int phase) {  // one thread: push a point to a path
    const int n = S.n[path];
    if (n >= PATH_TAIL) {
      for (int k = 0; k < PATH_TAIL - 1; k++) for (int j = 0; j < 3; j++) S.tail[path][k][j] = S.tail[path][k + 1][j];
      for (int j = 0; j < 3; j++) S.tail[path][PATH_TAIL - 1][j] = pt[j];
    } else {
      for (int j = 0; j < 3; j++) S.tail[path][n][j] = pt[j];
    }
    if (n < 0x3fffffff) S.n[path] = n + 1;
    double* o = out_paths + ((size_t)path * p.n_frames + f) * 3;
    o[0] = pt[0]; o[1] = pt[1]; o[2] = pt[2];
    if (out_assign) {
      int8_t* dst = out_assign + ((size_t)path * p.n_frames + f) * C;
      if (C == 8) {  // the 8 nibbles spread to 8 bytes, one store (rows of 8 bytes are 8-byte aligned)
        u64 x = comb & 0xffffffffull;
        x = (x | (x << 16)) & 0x0000ffff0000ffffull;
        x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
        x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
        *reinterpret_cast<u64*>(dst) = x;
      } else {
        for (int c = 0; c < C; c++) dst[c] = (int8_t)((comb >> (4 * c)) & 15);
      }
    }
    if (out_phase) out_phase[(size_t)path * p.n_frames + f] = (uint8_t)phase;
  };

  auto stage = [&](int buf, int L, long long off) {  // cp.async of one frame's candidate list (8-byte pieces: off is arbitrary)
    unsigned char* base = link_dyn + (size_t)buf * LINK_BUF_BYTES;
    u64* d_comb = reinterpret_cast<u64*>(base);
    double* d_err = reinterpret_cast<double*>(base + sizeof(u64) * LINK_STAGE_LEAVES);
    double* d_xyz = reinterpret_cast<double*>(base + (sizeof(u64) + sizeof(double)) * LINK_STAGE_LEAVES);
    if (L <= LINK_STAGE_LEAVES) {
      for (int i = tid; i < L; i += LINK_THREADS) { cp_async8(d_comb + i, leaf_comb + off + i); cp_async8(d_err + i, leaf_err + off + i); }
      for (int i = tid; i < 3 * L; i += LINK_THREADS) cp_async8(d_xyz + i, leaf_xyz + 3 * off + i);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // software pipeline over the frames: candidate-list extents two frames ahead, the list itself and the
  // detections one frame ahead
  const int nf = p.f1 - p.f0;
  int L_next = nf > 0 ? leaf_cnt[0] : 0, L_next2 = nf > 1 ? leaf_cnt[1] : 0;
  long long off_next = nf > 0 ? leaf_off[0] : 0, off_next2 = nf > 1 ? leaf_off[1] : 0;
  const int det_c = tid / TRI_MAX_DETS, det_d = tid % TRI_MAX_DETS;  // thread tid < C * TRI_MAX_DETS owns detection slot (c, d)
  const bool det_thread = tid < C * TRI_MAX_DETS;
  bool det_has_next = false;
  double det_x_next = 0, det_y_next = 0;
  auto fetch_det = [&](int f, bool& has, double& x, double& y) {
    has = false;
    if (det_thread && f < p.f1) {
      const int a = offs[(size_t)det_c * (p.n_frames + 1) + f], b = offs[(size_t)det_c * (p.n_frames + 1) + f + 1];
      has = det_d < b - a;
      if (has) { x = dets[2 * (size_t)(a + det_d)]; y = dets[2 * (size_t)(a + det_d) + 1]; }
    }
  };
  fetch_det(p.f0, det_has_next, det_x_next, det_y_next);
  stage(0, L_next, off_next);

  for (int f = p.f0; f < p.f1; f++) {
    const int k = f - p.f0;
    const int L = L_next;
    const long long off = off_next;
    L_next = L_next2; off_next = off_next2;
    const bool det_has = det_has_next;
    const double det_x = det_x_next, det_y = det_y_next;
    const bool staged = L <= LINK_STAGE_LEAVES;
    const unsigned char* cur = link_dyn + (size_t)(k & 1) * LINK_BUF_BYTES;
    const u64* lc = staged ? reinterpret_cast<const u64*>(cur) : leaf_comb + off;
    const double* le = staged ? reinterpret_cast<const double*>(cur + sizeof(u64) * LINK_STAGE_LEAVES) : leaf_err + off;
    const double* lx = staged ? reinterpret_cast<const double*>(cur + (sizeof(u64) + sizeof(double)) * LINK_STAGE_LEAVES) : leaf_xyz + 3 * off;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // this frame's list has landed; everyone is done with the previous frame (and its buffer)
    if (k + 1 < nf) stage((k + 1) & 1, L_next, off_next); else asm volatile("cp.async.commit_group;" ::: "memory");
    if (k + 2 < nf) { L_next2 = leaf_cnt[k + 2]; off_next2 = leaf_off[k + 2]; }
    fetch_det(f + 1, det_has_next, det_x_next, det_y_next);

    // ---- phase 1: tracking (:119-135) ----
    // (i) the pixel rays of this frame's detections, once (they do not depend on the path)
    if (det_thread) {
      s_ndet[tid] = det_has;
      if (det_has) ref::make_dir(ray, det_c, det_x, det_y, s_dir[tid]);
    }
    for (int i = tid; i < C * 16; i += LINK_THREADS) s_gate[i / 16][i % 16] = (i % 16) == 0 ? 0xffffffffu : 0u;  // choice 0 ("none") is always available
    if (tid < D) {
      const int n = S.n[tid];
      const double* last = S.tail[tid][min(max(n, 1), PATH_TAIL) - 1];
      s_active[tid] = n != 0 && !(last[0] == 0 && last[1] == 0 && last[2] == 0);  // :121-123
      s_cand[tid] = 0x7fffffff;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned m = 0;
      for (int np = 0; np < D; np++) m |= (s_active[np] ? 1u : 0u) << np;
      s_active_mask = m;
      s_found_mask = 0;
    }
    // (ii) the MAX_STEP ray gate of every (path, detection), :228-236 -- a path's last point is last frame's
    for (int i = tid; i < D * C * TRI_MAX_DETS; i += LINK_THREADS) {
      const int np = i / (C * TRI_MAX_DETS), k = i % (C * TRI_MAX_DETS), c = k / TRI_MAX_DETS, d = k % TRI_MAX_DETS;
      const int n = S.n[np];
      if (n == 0 || !s_ndet[k]) continue;
      const double* last = S.tail[np][min(n, PATH_TAIL) - 1];
      if (ref::dist_to_ray(ray.pos[c], s_dir[k], last[0], last[1], last[2]) < MAX_STEP) atomicOr(&s_gate[c][d + 1], 1u << np);
    }
    __syncthreads();
    // (iii) The leaves are stored in priority order, so "the first element the reference's priority_queue pops
    // that passes :241-246" is, per path, the admissible leaf of SMALLEST index.  All threads scan the list 256
    // leaves at a time; one AND over the cameras of the transposed gate words gives the set of paths a leaf is
    // gated for, the distance test runs only for those, and the winners are taken with atomicMin.  The scan
    // stops as soon as every tracked path has a candidate.  Combinations used by earlier paths are ignored here ...
    const int lane = tid & 31, warp = tid >> 5;
    auto gate_ok = [&](int np, u64 comb) {
      bool ok = true;
      for (int c = 0; c < C; c++) ok = ok && ((s_gate[c][(comb >> (4 * c)) & 15] >> np) & 1u);
      return ok;
    };
    {
      const unsigned active_mask = s_active_mask;
      for (int base = 0; active_mask && base < L; base += LINK_THREADS) {
        const int i = base + tid;
        if (i < L) {
          const u64 comb = lc[i];
          unsigned ap = active_mask;
          for (int c = 0; c < C; c++) ap &= s_gate[c][(comb >> (4 * c)) & 15];
          if (ap && le[i] < p.error_) {
            while (ap) {
              const int np = __ffs(ap) - 1;
              ap &= ap - 1;
              const double* last = S.tail[np][min(S.n[np], PATH_TAIL) - 1];
              if (dist3(lx + 3 * i, last) < MAX_STEP) {  // cv::norm(c.point - pos) < MAX_STEP, :244
                atomicMin(&s_cand[np], i);
                atomicOr(&s_found_mask, 1u << np);
              }
            }
          }
        }
        __syncthreads();
        if (__syncthreads_and(s_found_mask == active_mask)) break;  // (second barrier: everyone has read the mask before the next round writes it)
      }
    }
    if (tid < D && s_cand[tid] == 0x7fffffff) s_cand[tid] = -1;
    __syncthreads();
    // (iv) ... and resolved here in path order by warp 0: a candidate that does not collide with an earlier
    // path's pick is also the first of the filtered list; otherwise walk the list again with the filter.
    // Phase 2 (pickBestCombinations, :200-217) follows on the same warp: ONE pass in priority order keeping
    // every leaf that collides with nothing kept so far -- literally the reference's pop loop.
    if (warp == 0) {
      // lane np owns path np's speculative pick; the picks are confirmed in path order with shuffles (the
      // common case touches no memory), and all confirmed paths are written at once, one lane per path
      int cand = (lane < D && s_active[lane]) ? s_cand[lane] : -1;
      u64 comb = cand >= 0 ? lc[cand] : 0ull;
      bool clashed = false;  // my pick collides with a confirmed earlier one
      unsigned accepted = 0;
      for (int np = 0; np < D; np++) {
        int c_np = __shfl_sync(0xffffffffu, cand, np);
        if (c_np < 0) continue;
        if (__shfl_sync(0xffffffffu, (int)clashed, np)) {  // rare (404 of 14 738 picks on S09_D6): walk the list again with the used filter
          const int n_used = __popc(accepted);
          const double* last = S.tail[np][min(S.n[np], PATH_TAIL) - 1];
          c_np = -1;
          for (int base = 0; base < L && c_np < 0; base += 32) {
            const int i = base + lane;
            bool ok = i < L && gate_ok(np, lc[i]) && le[i] < p.error_;
            for (int u = 0; u < n_used && ok; u++) ok = !conflicts(lc[i], s_used[u]);
            if (ok) ok = dist3(lx + 3 * i, last) < MAX_STEP;
            const unsigned hit = __ballot_sync(0xffffffffu, ok);
            if (hit) c_np = base + __ffs(hit) - 1;
          }
          if (lane == np) { cand = c_np; comb = c_np >= 0 ? lc[c_np] : 0ull; }
          if (c_np < 0) continue;
        }
        const u64 pick = __shfl_sync(0xffffffffu, comb, np);
        if (lane == np) s_used[__popc(accepted)] = pick;
        accepted |= 1u << np;
        if (lane > np && cand >= 0) clashed = clashed || conflicts(comb, pick);
        __syncwarp();
      }
      if (accepted >> lane & 1u) emit(lane, f, comb, lx + 3 * cand, 1);
      if (lane == 0) { n_phase1 += __popc(accepted); s_processed = accepted; s_n_used = __popc(accepted); s_n_fin = 0; }
    }
    __syncthreads();
    if (__popc(s_processed) == D) continue;  // :137
    // ---- phase 2: pickBestCombinations (:200-217).  The reference pops the whole queue in priority order
    // and keeps every combination that collides with nothing kept so far; with the leaves in priority
    // order that is: repeatedly take the FIRST live leaf and kill everything that collides with it.  All
    // threads filter (each owns leaves tid, tid+256, ...; live flags in a register), one block-wide
    // min-index reduction per kept combination.
    {
      const int n_used = s_n_used;
      u64 live = 0;  // bit k <=> leaf tid + k * LINK_THREADS is still a candidate (L <= 64 * LINK_THREADS, checked in (A))
      for (int k = 0, i = tid; i < L; k++, i += LINK_THREADS) {
        bool ok = le[i] < p.error_;
        for (int u = 0; u < n_used && ok; u++) ok = !conflicts(lc[i], s_used[u]);
        live |= (ok ? 1ull : 0ull) << k;
      }
      for (int round = 0; round < LINK_MAX_FINAL; round++) {
        int first = live ? tid + (__ffsll((long long)live) - 1) * LINK_THREADS : 0x7fffffff;
        for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
        if (lane == 0) s_first[round & 1][warp] = first;
        __syncthreads();
        first = s_first[round & 1][0];
        for (int w = 1; w < LINK_THREADS / 32; w++) first = min(first, s_first[round & 1][w]);
        if (first == 0x7fffffff) break;
        const u64 pick = lc[first];
        if (tid == 0) { s_fin[round] = pick; s_fin_idx[round] = first; s_n_fin = round + 1; }
        for (int k = 0, i = tid; i < L; k++, i += LINK_THREADS)
          if ((live >> k & 1ull) && (i == first || conflicts(lc[i], pick))) live &= ~(1ull << k);
      }
    }
    __syncthreads();

    // ---- classifyPaths (:262-332): the (combination, path) tail distances in parallel, the rest on one thread ----
    {
      const int nf = s_n_fin;
      for (int i = tid; i < nf * D; i += LINK_THREADS) {
        const int ci = i / D, j = i % D;
        const int npc = min(S.n[j], PATH_TAIL);
        double dist = -1;
        if (npc > 0) {
          const double* pt = lx + 3 * s_fin_idx[ci];
          dist = 0;
          for (int t = 0; t < npc; t++) dist += dist3(S.tail[j][t], pt);
          dist /= (double)npc;
        }
        s_pdist[ci][j] = dist;
      }
    }
    __syncthreads();
    if (tid == 0) {
      const int nf = s_n_fin;
      int cp_comb[LINK_MAX_FINAL], cp_path[LINK_MAX_FINAL];
      double cp_err[LINK_MAX_FINAL];
      unsigned processed = s_processed;
      for (int i = 0; i < nf; i++) {
        int bestPath = 0;
        double bestDist = -1;
        for (int j = 0; j < D; j++) {
          if (processed >> j & 1u) continue;
          if (min(S.n[j], PATH_TAIL) == 0) continue;
          const double dist = s_pdist[i][j];
          if (dist < bestDist || bestDist == -1) { bestDist = dist; bestPath = j; }
        }
        // std::sort(greater<>) of <= 16 elements is an insertion sort in libstdc++: stable, ascending error
        int k = i - 1;
        while (k >= 0 && bestDist < cp_err[k]) { cp_comb[k + 1] = cp_comb[k]; cp_path[k + 1] = cp_path[k]; cp_err[k + 1] = cp_err[k]; k--; }
        cp_comb[k + 1] = i; cp_path[k + 1] = bestPath; cp_err[k + 1] = bestDist;
      }
      for (int k = 0; k < nf; k++) {
        int target = -1;
        if (processed >> cp_path[k] & 1u) {
          for (int i = 0; i < D; i++) if (S.n[i] == 0) { target = i; break; }
        } else {
          target = cp_path[k];
        }
        if (target != -1) {
          emit(target, f, s_fin[cp_comb[k]], lx + 3 * s_fin_idx[cp_comb[k]], 2);
          processed |= 1u << target;
          n_phase2++;
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  for (int i = tid; i < (int)(sizeof(LinkState) / sizeof(int)); i += LINK_THREADS) ((int*)state)[i] = ((const int*)&S)[i];
  if (tid == 0) { atomicAdd(&ctr->phase1, n_phase1); atomicAdd(&ctr->phase2, n_phase2); }
}

// Grow-only device buffer: the classifier's work space lives on the engine across calls (cudaMalloc /
// cudaFree of a few hundred MB per call cost up to a second, far more than the kernels).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) {
    if (bytes <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e == cudaSuccess) cap = bytes ? bytes : 1;
    return e;
  }
  template <typename T> T* as() { return static_cast<T*>(p); }
};

struct ClsWork {
  DevBuf offs, dets, paths, assign, phase, state, ctr, front, txyz, terr, lcomb, lerr, lxyz, loff, lcnt;
  // tri_classify_begin / tri_classify_finish: the enumerated shard waiting for its linking pass
  bool pending = false;
  ClsParams job;
  int job_max_frontier = 0;
};
static void free_cls_work(void* w) { delete static_cast<ClsWork*>(w); }

}  // namespace tri

using namespace tri;

#define TRI_CUDA(call)                                         \
  do {                                                         \
    cudaError_t err__ = (call);                                \
    if (err__ != cudaSuccess) return cuda_fail(err__, #call);  \
  } while (0)

extern "C" int tri_classify(tri_engine* e, int mode, unsigned flags, int n_drones, const int32_t* det_offsets,
                            const double* dets_xy, int n_frames, double* out_paths, int8_t* out_assign, uint8_t* out_phase,
                            tri_classify_stats* stats) {
  if (!e) return fail(TRI_ERR_ARG, "null engine");
  if (mode != TRI_MATRIX && mode != TRI_RAY) return fail(TRI_ERR_ARG, "mode must be TRI_MATRIX or TRI_RAY");
  const int C = e->n_cams;
  if (C > CLS_MAX_CAMS) return fail(TRI_ERR_ARG, "the classifier handles at most 16 cameras (the search is exponential in the camera count)");
  if (n_drones < 1 || n_drones > TRI_MAX_DRONES) return fail(TRI_ERR_ARG, "n_drones must be in [1, TRI_MAX_DRONES]");
  if (n_frames < 0 || !out_paths) return fail(TRI_ERR_ARG, "bad output arguments");
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n_frames == 0) return TRI_OK;
  if (!det_offsets) return fail(TRI_ERR_ARG, "null detection offsets");
  const size_t n_offs = (size_t)C * (n_frames + 1);
  int64_t n_det = 0;
  for (int c = 0; c < C; c++) {
    const int32_t* o = det_offsets + (size_t)c * (n_frames + 1);
    for (int f = 0; f < n_frames; f++) {
      if (o[f + 1] < o[f]) return fail(TRI_ERR_ARG, "detection offsets must be non-decreasing");
      if (o[f + 1] - o[f] > TRI_MAX_DETS) return fail(TRI_ERR_CAPACITY, "more than TRI_MAX_DETS detections on one camera in one frame");
    }
    n_det = std::max<int64_t>(n_det, o[n_frames]);
  }
  if (n_det > 0 && !dets_xy) return fail(TRI_ERR_ARG, "null detections");
  DeviceGuard g(e->device);
  cudaStream_t s = e->stream;

  ClsParams p;
  p.n_cams = C; p.n_drones = n_drones; p.n_frames = n_frames;
  p.solver = mode == TRI_MATRIX ? 0 : (flags & TRI_RAY_REFERENCE_LM) ? 1 : 2;
  p.error_ = mode == TRI_MATRIX ? MAX_ERROR_MATRIX : MAX_ERROR_RAY;  // DroneClassifier.cpp:3-10

  if (!e->cls_work) { e->cls_work = new ClsWork(); e->cls_work_free = free_cls_work; }
  ClsWork& W = *static_cast<ClsWork*>(e->cls_work);
  DevBuf &d_offs = W.offs, &d_dets = W.dets, &d_paths = W.paths, &d_assign = W.assign, &d_phase = W.phase, &d_state = W.state,
         &d_ctr = W.ctr, &d_front = W.front, &d_txyz = W.txyz, &d_terr = W.terr, &d_lcomb = W.lcomb, &d_lerr = W.lerr,
         &d_lxyz = W.lxyz, &d_loff = W.loff, &d_lcnt = W.lcnt;
  const size_t sz_paths = sizeof(double) * 3 * n_drones * (size_t)n_frames, sz_assign = (size_t)n_drones * n_frames * C,
               sz_phase = (size_t)n_drones * n_frames;
  TRI_CUDA(d_offs.alloc(sizeof(int32_t) * n_offs));
  TRI_CUDA(d_dets.alloc(sizeof(double) * 2 * n_det));
  TRI_CUDA(d_paths.alloc(sz_paths));
  TRI_CUDA(d_assign.alloc(sz_assign));
  TRI_CUDA(d_phase.alloc(sz_phase));
  TRI_CUDA(d_state.alloc(sizeof(LinkState)));
  TRI_CUDA(d_ctr.alloc(sizeof(ClsCounters)));
  TRI_CUDA(cudaMemcpyAsync(d_offs.p, det_offsets, sizeof(int32_t) * n_offs, cudaMemcpyHostToDevice, s));
  if (n_det) TRI_CUDA(cudaMemcpyAsync(d_dets.p, dets_xy, sizeof(double) * 2 * n_det, cudaMemcpyHostToDevice, s));
  TRI_CUDA(cudaMemsetAsync(d_paths.p, 0, sz_paths, s));
  TRI_CUDA(cudaMemsetAsync(d_assign.p, 0xff, sz_assign, s));
  TRI_CUDA(cudaMemsetAsync(d_phase.p, 0, sz_phase, s));
  TRI_CUDA(cudaMemsetAsync(d_state.p, 0, sizeof(LinkState), s));
  TRI_CUDA(cudaMemsetAsync(d_ctr.p, 0, sizeof(ClsCounters), s));

  int batch = std::min(n_frames, 8192);
  int cap = 1 << 14;
  long long leaf_cap = 4ll << 20;
  int grid = 0;
  auto alloc_work = [&]() -> int {
    // every resident CTA slot gets a frame: the tree expansion is latency-bound (dependent FP64 chains, idle
    // lanes on narrow levels), so occupancy is what hides it -- 4 CTAs per SM instead of 2: S09_D6 with the
    // exact LM 3.88 -> 2.82 s (profiles/r1_cls_grid_sweep.log)
    int per_sm = 2;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, enumerate_kernel, CLS_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
    if (const char* v = getenv("TRI_CLS_CTAS_PER_SM")) per_sm = std::max(1, atoi(v));  // tuning override
    grid = std::max(1, std::min(batch, per_sm * e->sm_count));
    TRI_CUDA(d_front.alloc(sizeof(u64) * 2 * (size_t)cap * grid));
    TRI_CUDA(d_txyz.alloc(sizeof(double) * 3 * (size_t)cap * grid));
    TRI_CUDA(d_terr.alloc(sizeof(double) * (size_t)cap * grid));
    TRI_CUDA(d_lcomb.alloc(sizeof(u64) * leaf_cap));
    TRI_CUDA(d_lerr.alloc(sizeof(double) * leaf_cap));
    TRI_CUDA(d_lxyz.alloc(sizeof(double) * 3 * leaf_cap));
    TRI_CUDA(d_loff.alloc(sizeof(long long) * batch));
    TRI_CUDA(d_lcnt.alloc(sizeof(int) * batch));
    return TRI_OK;
  };
  int st = alloc_work();
  if (st != TRI_OK) return st;
  constexpr int LINK_DYN_BYTES = 2 * (int)((sizeof(u64) + 4 * sizeof(double)) * LINK_STAGE_LEAVES);
  TRI_CUDA(cudaFuncSetAttribute(link_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LINK_DYN_BYTES));

  ClsCounters h{};
  int max_frontier = 0;
  for (int f0 = 0; f0 < n_frames;) {
    const int f1 = std::min(n_frames, f0 + batch);
    p.f0 = f0; p.f1 = f1; p.cap = cap; p.leaf_cap = leaf_cap;
    ClsCounters before;
    TRI_CUDA(cudaMemcpyAsync(&before, d_ctr.p, sizeof(before), cudaMemcpyDeviceToHost, s));
    TRI_CUDA(cudaStreamSynchronize(s));
    TRI_CUDA(cudaMemsetAsync(&d_ctr.as<ClsCounters>()->leaf_total, 0, sizeof(u64), s));
    enumerate_kernel<<<grid, CLS_THREADS, 0, s>>>(e->rig64, e->ray, p, d_offs.as<int32_t>(), d_dets.as<double>(), d_front.as<u64>(),
                                                   d_txyz.as<double>(), d_terr.as<double>(), d_lcomb.as<u64>(), d_lerr.as<double>(),
                                                   d_lxyz.as<double>(), d_loff.as<long long>(), d_lcnt.as<int>(), d_ctr.as<ClsCounters>());
    e->launches++;
    TRI_CUDA(cudaGetLastError());
    TRI_CUDA(cudaMemcpyAsync(&h, d_ctr.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    TRI_CUDA(cudaStreamSynchronize(s));
    if (h.bad_input == 2) return fail(TRI_ERR_CAPACITY, "more than 16384 candidate combinations in one frame");
    if (h.bad_input) return fail(TRI_ERR_ARG, "malformed detection offsets");
    if (h.overflow_frontier || h.overflow_leaves) {
      // grow the work buffers (or shrink the batch) and redo this batch; the statistics of the
      // aborted attempt are rolled back
      if (h.overflow_frontier) { if (cap >= (1 << 22)) return fail(TRI_ERR_CAPACITY, "combination frontier exceeds 4M nodes in one frame"); cap *= 4; batch = std::max(1, batch / 4); }
      if (h.overflow_leaves) { if (batch > 64) batch /= 4; else if (leaf_cap < (256ll << 20)) leaf_cap *= 4; else return fail(TRI_ERR_CAPACITY, "candidate list exceeds device buffer"); }
      before.overflow_frontier = before.overflow_leaves = 0;
      TRI_CUDA(cudaMemcpyAsync(d_ctr.p, &before, sizeof(before), cudaMemcpyHostToDevice, s));
      TRI_CUDA(cudaStreamSynchronize(s));
      if ((st = alloc_work()) != TRI_OK) return st;
      continue;
    }
    max_frontier = std::max(max_frontier, h.max_frontier);
    link_kernel<<<1, LINK_THREADS, LINK_DYN_BYTES, s>>>(e->ray, p, d_offs.as<int32_t>(), d_dets.as<double>(), d_lcomb.as<u64>(), d_lerr.as<double>(),
                                            d_lxyz.as<double>(), d_loff.as<long long>(), d_lcnt.as<int>(), d_state.as<LinkState>(),
                                            d_paths.as<double>(), out_assign ? d_assign.as<int8_t>() : nullptr,
                                            out_phase ? d_phase.as<uint8_t>() : nullptr, d_ctr.as<ClsCounters>());
    e->launches++;
    TRI_CUDA(cudaGetLastError());
    f0 = f1;
  }
  TRI_CUDA(cudaMemcpyAsync(out_paths, d_paths.p, sz_paths, cudaMemcpyDeviceToHost, s));
  if (out_assign) TRI_CUDA(cudaMemcpyAsync(out_assign, d_assign.p, sz_assign, cudaMemcpyDeviceToHost, s));
  if (out_phase) TRI_CUDA(cudaMemcpyAsync(out_phase, d_phase.p, sz_phase, cudaMemcpyDeviceToHost, s));
  TRI_CUDA(cudaMemcpyAsync(&h, d_ctr.p, sizeof(h), cudaMemcpyDeviceToHost, s));
  TRI_CUDA(cudaStreamSynchronize(s));
  if (stats) {
    stats->nodes = (int64_t)h.nodes; stats->solves = (int64_t)h.solves; stats->leaves = (int64_t)h.leaves;
    stats->lm_iters = (int64_t)h.lm_iters; stats->phase1 = (int64_t)h.phase1; stats->phase2 = (int64_t)h.phase2;
    stats->ties = (int64_t)h.ties; stats->max_frontier = max_frontier;
  }
  return TRI_OK;
}

// ---- frame-sharded classification: candidate generation is independent per frame, linking is a
// chain.  tri_classify_begin enumerates a contiguous shard of the sequence and keeps every frame's candidate
// list on the device; tri_classify_finish links the shard starting from the tracking state the previous shard
// ended with (n_drones x (last point + 3-point tail) + counters, an opaque tri_classify_state_bytes() blob) and
// returns the state for the next one.  All ranks enumerate at once; only the small state travels along the chain.
extern "C" int tri_classify_state_bytes(void) { return (int)sizeof(LinkState); }

extern "C" int tri_classify_begin(tri_engine* e, int mode, unsigned flags, int n_drones, const int32_t* det_offsets,
                                  const double* dets_xy, int n_frames) {
  if (!e) return fail(TRI_ERR_ARG, "null engine");
  if (mode != TRI_MATRIX && mode != TRI_RAY) return fail(TRI_ERR_ARG, "mode must be TRI_MATRIX or TRI_RAY");
  const int C = e->n_cams;
  if (C > CLS_MAX_CAMS) return fail(TRI_ERR_ARG, "the classifier handles at most 16 cameras (the search is exponential in the camera count)");
  if (n_drones < 1 || n_drones > TRI_MAX_DRONES) return fail(TRI_ERR_ARG, "n_drones must be in [1, TRI_MAX_DRONES]");
  if (n_frames < 0) return fail(TRI_ERR_ARG, "bad frame count");
  if (n_frames > 0 && !det_offsets) return fail(TRI_ERR_ARG, "null detection offsets");
  const size_t n_offs = (size_t)C * (n_frames + 1);
  int64_t n_det = 0;
  for (int c = 0; c < C && n_frames > 0; c++) {
    const int32_t* o = det_offsets + (size_t)c * (n_frames + 1);
    for (int f = 0; f < n_frames; f++) {
      if (o[f + 1] < o[f]) return fail(TRI_ERR_ARG, "detection offsets must be non-decreasing");
      if (o[f + 1] - o[f] > TRI_MAX_DETS) return fail(TRI_ERR_CAPACITY, "more than TRI_MAX_DETS detections on one camera in one frame");
    }
    n_det = std::max<int64_t>(n_det, o[n_frames]);
  }
  if (n_det > 0 && !dets_xy) return fail(TRI_ERR_ARG, "null detections");
  DeviceGuard g(e->device);
  cudaStream_t s = e->stream;
  if (!e->cls_work) { e->cls_work = new ClsWork(); e->cls_work_free = free_cls_work; }
  ClsWork& W = *static_cast<ClsWork*>(e->cls_work);
  W.pending = false;
  ClsParams p;
  p.n_cams = C; p.n_drones = n_drones; p.n_frames = n_frames;
  p.solver = mode == TRI_MATRIX ? 0 : (flags & TRI_RAY_REFERENCE_LM) ? 1 : 2;
  p.error_ = mode == TRI_MATRIX ? MAX_ERROR_MATRIX : MAX_ERROR_RAY;
  p.f0 = 0; p.f1 = n_frames;
  W.job = p;
  W.job_max_frontier = 0;
  if (n_frames == 0) { W.pending = true; return TRI_OK; }
  TRI_CUDA(W.offs.alloc(sizeof(int32_t) * n_offs));
  TRI_CUDA(W.dets.alloc(sizeof(double) * 2 * n_det));
  TRI_CUDA(W.ctr.alloc(sizeof(ClsCounters)));
  TRI_CUDA(W.loff.alloc(sizeof(long long) * n_frames));
  TRI_CUDA(W.lcnt.alloc(sizeof(int) * n_frames));
  TRI_CUDA(cudaMemcpyAsync(W.offs.p, det_offsets, sizeof(int32_t) * n_offs, cudaMemcpyHostToDevice, s));
  if (n_det) TRI_CUDA(cudaMemcpyAsync(W.dets.p, dets_xy, sizeof(double) * 2 * n_det, cudaMemcpyHostToDevice, s));
  int per_sm = 2;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, enumerate_kernel, CLS_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
  int cap = 1 << 14;
  long long leaf_cap = std::max<long long>(4ll << 20, 1024ll * n_frames);
  for (;;) {  // the whole shard's candidates stay resident: on overflow grow the buffers and enumerate again
    const int grid = std::max(1, std::min(n_frames, per_sm * e->sm_count));
    TRI_CUDA(W.front.alloc(sizeof(u64) * 2 * (size_t)cap * grid));
    TRI_CUDA(W.txyz.alloc(sizeof(double) * 3 * (size_t)cap * grid));
    TRI_CUDA(W.terr.alloc(sizeof(double) * (size_t)cap * grid));
    TRI_CUDA(W.lcomb.alloc(sizeof(u64) * leaf_cap));
    TRI_CUDA(W.lerr.alloc(sizeof(double) * leaf_cap));
    TRI_CUDA(W.lxyz.alloc(sizeof(double) * 3 * leaf_cap));
    TRI_CUDA(cudaMemsetAsync(W.ctr.p, 0, sizeof(ClsCounters), s));
    p.cap = cap; p.leaf_cap = leaf_cap;
    enumerate_kernel<<<grid, CLS_THREADS, 0, s>>>(e->rig64, e->ray, p, W.offs.as<int32_t>(), W.dets.as<double>(), W.front.as<u64>(),
                                                   W.txyz.as<double>(), W.terr.as<double>(), W.lcomb.as<u64>(), W.lerr.as<double>(),
                                                   W.lxyz.as<double>(), W.loff.as<long long>(), W.lcnt.as<int>(), W.ctr.as<ClsCounters>());
    e->launches++;
    TRI_CUDA(cudaGetLastError());
    ClsCounters h{};
    TRI_CUDA(cudaMemcpyAsync(&h, W.ctr.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    TRI_CUDA(cudaStreamSynchronize(s));
    if (h.bad_input == 2) return fail(TRI_ERR_CAPACITY, "more than 16384 candidate combinations in one frame");
    if (h.bad_input) return fail(TRI_ERR_ARG, "malformed detection offsets");
    if (h.overflow_frontier) { if (cap >= (1 << 22)) return fail(TRI_ERR_CAPACITY, "combination frontier exceeds 4M nodes in one frame"); cap *= 4; continue; }
    if (h.overflow_leaves) { if (leaf_cap >= (256ll << 20)) return fail(TRI_ERR_CAPACITY, "candidate list exceeds device buffer"); leaf_cap *= 4; continue; }
    W.job_max_frontier = h.max_frontier;
    break;
  }
  W.job = p;
  W.pending = true;
  return TRI_OK;
}

extern "C" int tri_classify_finish(tri_engine* e, const void* state_in, void* state_out, double* out_paths, int8_t* out_assign,
                                   uint8_t* out_phase, tri_classify_stats* stats) {
  if (!e || !e->cls_work) return fail(TRI_ERR_ARG, "tri_classify_finish without tri_classify_begin");
  ClsWork& W = *static_cast<ClsWork*>(e->cls_work);
  if (!W.pending) return fail(TRI_ERR_ARG, "tri_classify_finish without tri_classify_begin");
  W.pending = false;
  const ClsParams p = W.job;
  const int C = p.n_cams, n_frames = p.n_frames, n_drones = p.n_drones;
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n_frames == 0) {
    if (state_out) { if (state_in) memcpy(state_out, state_in, sizeof(LinkState)); else memset(state_out, 0, sizeof(LinkState)); }
    return TRI_OK;
  }
  if (!out_paths) return fail(TRI_ERR_ARG, "bad output arguments");
  DeviceGuard g(e->device);
  cudaStream_t s = e->stream;
  const size_t sz_paths = sizeof(double) * 3 * n_drones * (size_t)n_frames, sz_assign = (size_t)n_drones * n_frames * C,
               sz_phase = (size_t)n_drones * n_frames;
  TRI_CUDA(W.paths.alloc(sz_paths));
  TRI_CUDA(W.assign.alloc(sz_assign));
  TRI_CUDA(W.phase.alloc(sz_phase));
  TRI_CUDA(W.state.alloc(sizeof(LinkState)));
  TRI_CUDA(cudaMemsetAsync(W.paths.p, 0, sz_paths, s));
  TRI_CUDA(cudaMemsetAsync(W.assign.p, 0xff, sz_assign, s));
  TRI_CUDA(cudaMemsetAsync(W.phase.p, 0, sz_phase, s));
  if (state_in) TRI_CUDA(cudaMemcpyAsync(W.state.p, state_in, sizeof(LinkState), cudaMemcpyHostToDevice, s));
  else TRI_CUDA(cudaMemsetAsync(W.state.p, 0, sizeof(LinkState), s));
  constexpr int LINK_DYN_BYTES = 2 * (int)((sizeof(u64) + 4 * sizeof(double)) * LINK_STAGE_LEAVES);
  TRI_CUDA(cudaFuncSetAttribute(link_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LINK_DYN_BYTES));
  link_kernel<<<1, LINK_THREADS, LINK_DYN_BYTES, s>>>(e->ray, p, W.offs.as<int32_t>(), W.dets.as<double>(), W.lcomb.as<u64>(), W.lerr.as<double>(),
                                                      W.lxyz.as<double>(), W.loff.as<long long>(), W.lcnt.as<int>(), W.state.as<LinkState>(),
                                                      W.paths.as<double>(), out_assign ? W.assign.as<int8_t>() : nullptr,
                                                      out_phase ? W.phase.as<uint8_t>() : nullptr, W.ctr.as<ClsCounters>());
  e->launches++;
  TRI_CUDA(cudaGetLastError());
  ClsCounters h{};
  TRI_CUDA(cudaMemcpyAsync(out_paths, W.paths.p, sz_paths, cudaMemcpyDeviceToHost, s));
  if (out_assign) TRI_CUDA(cudaMemcpyAsync(out_assign, W.assign.p, sz_assign, cudaMemcpyDeviceToHost, s));
  if (out_phase) TRI_CUDA(cudaMemcpyAsync(out_phase, W.phase.p, sz_phase, cudaMemcpyDeviceToHost, s));
  if (state_out) TRI_CUDA(cudaMemcpyAsync(state_out, W.state.p, sizeof(LinkState), cudaMemcpyDeviceToHost, s));
  TRI_CUDA(cudaMemcpyAsync(&h, W.ctr.p, sizeof(h), cudaMemcpyDeviceToHost, s));
  TRI_CUDA(cudaStreamSynchronize(s));
  if (stats) {
    stats->nodes = (int64_t)h.nodes; stats->solves = (int64_t)h.solves; stats->leaves = (int64_t)h.leaves;
    stats->lm_iters = (int64_t)h.lm_iters; stats->phase1 = (int64_t)h.phase1; stats->phase2 = (int64_t)h.phase2;
    stats->ties = (int64_t)h.ties; stats->max_frontier = W.job_max_frontier;
  }
  return TRI_OK;
}

// One C++ process driving several GPUs: every engine enumerates its contiguous frame range on its own host thread,
// then the shards are linked in order on the caller's thread with the state handed along.  Same outputs as
// tri_classify on the whole sequence.
extern "C" int tri_classify_multi(tri_engine* const* engines, int n_engines, int mode, unsigned flags, int n_drones,
                                  const int32_t* det_offsets, const double* dets_xy, int n_frames, double* out_paths,
                                  int8_t* out_assign, uint8_t* out_phase, tri_classify_stats* stats) {
  if (!engines || n_engines < 1) return fail(TRI_ERR_ARG, "no engines");
  for (int g = 0; g < n_engines; g++)
    if (!engines[g] || engines[g]->n_cams != engines[0]->n_cams) return fail(TRI_ERR_ARG, "engines must hold the same rig");
  if (n_engines == 1) return tri_classify(engines[0], mode, flags, n_drones, det_offsets, dets_xy, n_frames, out_paths, out_assign, out_phase, stats);
  if (n_frames < 0 || !out_paths || (n_frames > 0 && !det_offsets)) return fail(TRI_ERR_ARG, "bad arguments");
  if (n_drones < 1 || n_drones > TRI_MAX_DRONES) return fail(TRI_ERR_ARG, "n_drones must be in [1, TRI_MAX_DRONES]");
  if (n_frames == 0) { if (stats) memset(stats, 0, sizeof(*stats)); return TRI_OK; }
  const int C = engines[0]->n_cams;
  auto cut = [&](int g) { return (int)((int64_t)n_frames * g / n_engines); };
  // per-shard CSR: offsets rebased to the shard, detections gathered camera by camera
  std::vector<std::vector<int32_t>> offs(n_engines);
  std::vector<std::vector<double>> dets(n_engines);
  for (int g = 0; g < n_engines; g++) {
    const int f0 = cut(g), f1 = cut(g + 1);
    offs[g].assign((size_t)C * (f1 - f0 + 1), 0);
    int32_t base = 0;
    for (int c = 0; c < C; c++) {
      const int32_t* o = det_offsets + (size_t)c * (n_frames + 1);
      const int32_t a = o[f0], b = o[f1];
      if (b < a) return fail(TRI_ERR_ARG, "detection offsets must be non-decreasing");
      for (int f = f0; f <= f1; f++) offs[g][(size_t)c * (f1 - f0 + 1) + (f - f0)] = o[f] - a + base;
      if (b > a) {
        if (!dets_xy) return fail(TRI_ERR_ARG, "null detections");
        dets[g].insert(dets[g].end(), dets_xy + 2 * (size_t)a, dets_xy + 2 * (size_t)b);
      }
      base += b - a;
    }
  }
  std::vector<int> status(n_engines, TRI_OK);
  std::vector<std::string> msg(n_engines);
  std::vector<std::thread> workers;
  for (int g = 0; g < n_engines; g++)
    workers.emplace_back([&, g]() {
      status[g] = tri_classify_begin(engines[g], mode, flags, n_drones, offs[g].data(), dets[g].empty() ? nullptr : dets[g].data(), cut(g + 1) - cut(g));
      if (status[g] != TRI_OK) msg[g] = tri_last_error();
    });
  for (std::thread& t : workers) t.join();
  for (int g = 0; g < n_engines; g++)
    if (status[g] != TRI_OK) return fail(status[g], msg[g]);
  if (stats) memset(stats, 0, sizeof(*stats));
  std::vector<unsigned char> state(sizeof(LinkState), 0);
  for (int g = 0; g < n_engines; g++) {
    const int f0 = cut(g), nf = cut(g + 1) - f0;
    std::vector<double> paths((size_t)3 * n_drones * nf);
    std::vector<int8_t> assign(out_assign ? (size_t)n_drones * nf * C : 0);
    std::vector<uint8_t> phase(out_phase ? (size_t)n_drones * nf : 0);
    tri_classify_stats st{};
    const int rc = tri_classify_finish(engines[g], g == 0 ? nullptr : state.data(), state.data(), nf ? paths.data() : out_paths,
                                       out_assign ? assign.data() : nullptr, out_phase ? phase.data() : nullptr, &st);
    if (rc != TRI_OK) return rc;
    for (int d = 0; d < n_drones; d++) {  // [drone][frame] rows of the shard into the whole sequence's rows
      if (nf) memcpy(out_paths + ((size_t)d * n_frames + f0) * 3, paths.data() + (size_t)d * nf * 3, sizeof(double) * 3 * nf);
      if (out_assign && nf) memcpy(out_assign + ((size_t)d * n_frames + f0) * C, assign.data() + (size_t)d * nf * C, (size_t)nf * C);
      if (out_phase && nf) memcpy(out_phase + (size_t)d * n_frames + f0, phase.data() + (size_t)d * nf, nf);
    }
    if (stats) {
      stats->nodes += st.nodes; stats->solves += st.solves; stats->leaves += st.leaves; stats->lm_iters += st.lm_iters;
      stats->phase1 += st.phase1; stats->phase2 += st.phase2; stats->ties += st.ties;
      stats->max_frontier = std::max(stats->max_frontier, st.max_frontier);
    }
  }
  return TRI_OK;
}
