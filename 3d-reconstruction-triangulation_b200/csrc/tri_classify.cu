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
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "tri_classify.cuh"

namespace tri {

// ---- (A) candidate generation ------------------------------------------------------------------
__global__ void __launch_bounds__(CLS_THREADS)
enumerate_kernel(const __grid_constant__ DltRig<double> dlt, const __grid_constant__ RayRig ray, ClsParams p,
                 const int32_t* __restrict__ offs, const double* __restrict__ dets, u64* __restrict__ front,
                 double* __restrict__ tmp_xyz, double* __restrict__ tmp_err, u64* __restrict__ leaf_rec,
                 long long* __restrict__ leaf_off, int* __restrict__ leaf_cnt, int* __restrict__ hdr,
                 FrameDet* __restrict__ fdet, long long* __restrict__ fdet_off, int* __restrict__ fdet_cnt, ClsCounters* ctr) {
  __shared__ int s_n[CLS_MAX_CAMS], s_pref[CLS_MAX_CAMS + 1], s_hist[CLS_MAX_CAMS + 2];
  __shared__ long long s_doff;
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
    __syncthreads();  // s_n is complete
    if (tid == 0) {
      buf0[0] = 0;
      int n = 0;
      for (int c = 0; c < C; c++) { s_pref[c] = n; n += s_n[c]; }
      s_pref[C] = n;
      s_doff = (long long)atomicAdd(&ctr->fdet_total, (u64)n);  // the buffer holds every detection of the batch: no overflow
      fdet_off[f - p.f0] = s_doff;
      fdet_cnt[f - p.f0] = n;
    }
    __syncthreads();
    // the frame's detections with their pixel rays, camera-major: the linking pass gates them against every path
    for (int i = tid; i < C * TRI_MAX_DETS; i += CLS_THREADS) {
      const int c = i / TRI_MAX_DETS, d = i % TRI_MAX_DETS;
      if (d < s_n[c]) {
        FrameDet fd;
        ref::make_dir(ray, c, s_px[c][d], s_py[c][d], fd.dir);
        fd.org[0] = ray.pos[c][0]; fd.org[1] = ray.pos[c][1]; fd.org[2] = ray.pos[c][2];
        for (int j = 0; j < 3; j++) { fd.dirf[j] = (float)fd.dir[j]; fd.orgf[j] = (float)fd.org[j]; }
        fd.cam = c; fd.slot = d;
        fdet[s_doff + s_pref[c] + d] = fd;
      }
    }
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
      u64* t = fin; fin = fout; fout = t;
      if (m == 0) break;
      if (last) {
        // publish this frame's leaves as records into a contiguous range of the global leaf array
        if (tid == 0) {
          long long off = (long long)atomicAdd(&ctr->leaf_total, (u64)m);
          if (off + m > p.leaf_cap) { atomicExch(&ctr->overflow_leaves, 1); off = -1; }
          s_off = off;
        }
        for (int i = tid; i < CLS_MAX_CAMS + 2; i += CLS_THREADS) s_hist[i] = 0;
        __syncthreads();
        const long long off = s_off;
        if (off >= 0) {
          // ... in PRIORITY order: rank of leaf i = number of leaves the reference's priority_queue pops
          // before it = those with fewer unused cameras, then smaller error (Combination::operator<,
          // :12-20), ties by DFS order.  Rank by counting (m is ~1e3; frames are independent, so this is
          // parallel work) and scatter; the sequential linking kernel then only walks prefixes.
          const u64 cam_bits = C == 16 ? ~0ull : ((1ull << (4 * C)) - 1);
          const int RW = rec_words(p.W);
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
            atomicAdd(&s_hist[zi + 1], 1);
            u64* rec = leaf_rec + (size_t)(off + rank) * RW;
            u64 mk[4] = {0, 0, 0, 0};
            for (int c = 0; c < C; c++) {
              const int k = (int)((ci >> (4 * c)) & 15);
              if (k) { const int bit = s_pref[c] + k - 1; mk[bit >> 6] |= 1ull << (bit & 63); }
            }
            if (!(ei < p.error_)) mk[p.W - 1] |= 1ull << 63;  // poison: a leaf (:185-196) that no acceptance test passes (:209, :243)
            for (int w = 0; w < p.W; w++) rec[w] = mk[w];
            rec[p.W] = (u64)__double_as_longlong(t_xyz[3 * i]);
            rec[p.W + 1] = (u64)__double_as_longlong(t_xyz[3 * i + 1]);
            rec[p.W + 2] = (u64)__double_as_longlong(t_xyz[3 * i + 2]);
            rec[p.W + 3] = ci;
          }
          if (n_tie) atomicAdd(&ctr->ties, n_tie);
          __syncthreads();
          if (tid == 0) {
            leaf_off[f - p.f0] = off; leaf_cnt[f - p.f0] = m; atomicAdd(&ctr->leaves, (u64)m);
            int run = 0;
            int* zs = hdr + (size_t)(f - p.f0) * HDR_INTS;
            for (int z = 0; z <= C + 1; z++) { run += s_hist[z]; zs[z] = run; }  // zs[z] = leaves with fewer than z unused cameras
            for (int c = 0; c <= C; c++) zs[HDR_PREF + c] = s_pref[c];
          }
        } else if (tid == 0) {
          leaf_off[f - p.f0] = 0; leaf_cnt[f - p.f0] = 0;
        }
        m = -1;  // done
      }
    }
    if (m >= 0 && tid == 0) {  // the tree died out (or no cameras)
      leaf_off[f - p.f0] = 0; leaf_cnt[f - p.f0] = 0;
      for (int z = 0; z <= C + 1; z++) hdr[(size_t)(f - p.f0) * HDR_INTS + z] = 0;
      for (int c = 0; c <= C; c++) hdr[(size_t)(f - p.f0) * HDR_INTS + HDR_PREF + c] = s_pref[c];
    }
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
// One CTA of eight warps per sequence, TWO block barriers per frame.  Everything that does not depend on the tracking
// state was prepared by (A): the leaves in priority order as (detection mask, point, combination) records, the number
// of leaves per count of unused cameras, the frame's detections with their pixel rays.  A frame's records, detections
// and header arrive in shared memory as three bulk async copies (cp.async.bulk -> UBLKCP) that complete on an
// mbarrier, issued one frame ahead into the other buffer.  Per frame:
//   speculative phase 1, one warp per tracked path, all paths at once:
//     gate    the MAX_STEP ray gate (:228-236) with lane <-> detection: the ballot of one 32-detection test IS a slice
//             of the path's gate mask.  The distance runs in single precision first and in the reference's
//             double-precision operation order only where single precision cannot decide.
//     scan    the FIRST leaf (priority order) whose mask lies inside the gate and whose point is within MAX_STEP of the
//             path's last point (:241-246), ignoring earlier paths' picks; it starts at the first leaf with at least
//             as many unused cameras as the gate leaves empty and tests 128 leaves per step.
//   warp 0: confirmation in path order with shuffles (a pick that collides with an earlier one -- 404 of 14 738 on
//     S09_D6 -- is scanned again with the used mask); phase 2, pickBestCombinations (:200-217), as ONE forward pass over
//     the list with a running used mask -- literally the reference's pop loop; classifyPaths (:262-332) with the tail
//     distances one (combination, open path) pair per lane and the ordered assignment by warp-wide minimum extraction.
// Round 1 ran this with ~15 block barriers per frame, the pixel rays and the whole phase-2 filter inside the sequential
// kernel (29 ms for S09_D6's 3000 frames, profiles/r1_link_kernel_lines.txt); now 20 ms and falling (profiles/r2_*).
__device__ __forceinline__ uint32_t cls_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cls_mbar_init(u64* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cls_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void cls_mbar_expect_tx(u64* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cls_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cls_mbar_wait(u64* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(cls_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cls_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(cls_smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(cls_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ double dist3(const double* a, const double* b) {  // cv::norm(a - b)
  const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
  return sqrt(x * x + y * y + z * z);
}
// sqrt(s) < MAX_STEP without the square root unless s is within rounding reach of MAX_STEP^2 (sqrt is monotonic and
// correctly rounded, so away from the boundary the two compares agree)
__device__ __forceinline__ bool sqrt_below_step(double s) {
  const double t = MAX_STEP * MAX_STEP;
  if (s < t * (1 - 1e-12)) return true;
  if (s > t * (1 + 1e-12)) return false;
  return sqrt(s) < MAX_STEP;
}

template <int W>
struct LinkLayout {
  static constexpr int RW = rec_words(W);
  static constexpr int MAX_LEAVES = W == 2 ? 1664 : 960;  // staged leaves per frame; longer lists are read from global memory
  static constexpr int REC_BYTES = MAX_LEAVES * RW * 8;
  static constexpr int DET_BYTES = LINK_MAX_DETS * (int)sizeof(FrameDet);
  static constexpr int HDR_BYTES = HDR_INTS * 4;
  static constexpr int BUF_BYTES = REC_BYTES + DET_BYTES + HDR_BYTES;
  static constexpr int BYTES = 2 * BUF_BYTES + 16;  // + the two mbarriers
};

constexpr int LINK_WARPS = 8;
constexpr int LINK_THREADS = 32 * LINK_WARPS;

template <int W>
__global__ void __launch_bounds__(LINK_THREADS)
link_kernel(const __grid_constant__ RayRig ray, ClsParams p, const int2* __restrict__ seq_bounds,
            const u64* __restrict__ leaf_rec, const long long* __restrict__ leaf_off, const int* __restrict__ leaf_cnt,
            const int* __restrict__ hdr, const FrameDet* __restrict__ fdet, const long long* __restrict__ fdet_off,
            const int* __restrict__ fdet_cnt, LinkState* state, double* __restrict__ out_paths, int8_t* __restrict__ out_assign,
            uint8_t* __restrict__ out_phase, ClsCounters* ctr) {
  using L_ = LinkLayout<W>;
  constexpr int RW = L_::RW;
  constexpr u64 POISON = 1ull << 63;  // in word W - 1
  extern __shared__ __align__(128) unsigned char link_dyn[];
  __shared__ LinkState S;
  __shared__ float s_lastf[TRI_MAX_DRONES][4];     // single-precision copy of each path's last point
  __shared__ unsigned s_gate[TRI_MAX_DRONES][2 * W];  // 32-bit slices of each tracked path's gate mask
  __shared__ int s_spec[TRI_MAX_DRONES];           // each tracked path's speculative pick (leaf index or -1)
  __shared__ int s_fin_idx[LINK_MAX_FINAL], s_cp_path[LINK_MAX_FINAL];
  __shared__ double s_cp_err[LINK_MAX_FINAL];
  __shared__ double s_pdist[LINK_MAX_FINAL][TRI_MAX_DRONES];
  u64* full = reinterpret_cast<u64*>(link_dyn + 2 * L_::BUF_BYTES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, C = p.n_cams, D = p.n_drones;
  const int fa = seq_bounds ? seq_bounds[blockIdx.x].x : p.f0, fb = seq_bounds ? seq_bounds[blockIdx.x].y : p.f1;
  LinkState* st = state + blockIdx.x;
  for (int i = tid; i < (int)(sizeof(LinkState) / sizeof(int)); i += LINK_THREADS) ((int*)&S)[i] = ((const int*)st)[i];
  if (tid == 0) {
    cls_mbar_init(&full[0], 1); cls_mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < D) {
    const double* last = S.tail[tid][min(max(S.n[tid], 1), PATH_TAIL) - 1];
    s_lastf[tid][0] = (float)last[0]; s_lastf[tid][1] = (float)last[1]; s_lastf[tid][2] = (float)last[2];
  }
  u64 n_phase1 = 0, n_phase2 = 0;  // thread 0
  bool overflow_final = false;
#ifdef TRI_TUNING
  u64 prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long prof_t = clock64();
#endif

  // stage frame f's records / detections / header into buffer b (thread 0)
  auto stage = [&](int b, int f, int L, long long off, int nd, long long doff) {
    unsigned char* base = link_dyn + (size_t)b * L_::BUF_BYTES;
    const uint32_t rec_bytes = L <= L_::MAX_LEAVES ? (uint32_t)L * RW * 8 : 0, det_bytes = (uint32_t)nd * (uint32_t)sizeof(FrameDet);
    cls_mbar_expect_tx(&full[b], rec_bytes + det_bytes + L_::HDR_BYTES);
    if (rec_bytes) cls_bulk_load(base, leaf_rec + (size_t)off * RW, rec_bytes, &full[b]);
    if (det_bytes) cls_bulk_load(base + L_::REC_BYTES, fdet + doff, det_bytes, &full[b]);
    cls_bulk_load(base + L_::REC_BYTES + L_::DET_BYTES, hdr + (size_t)(f - p.f0) * HDR_INTS, L_::HDR_BYTES, &full[b]);
  };
  auto emit = [&](int path, int f, const u64* r, int phase) {  // one thread: push leaf r's point to a path
    const double x = __longlong_as_double((long long)r[W]), y = __longlong_as_double((long long)r[W + 1]), z = __longlong_as_double((long long)r[W + 2]);
    const u64 comb = r[W + 3];
    const int n = S.n[path];
    double(*t)[3] = S.tail[path];
    if (n >= PATH_TAIL) {
      for (int k = 0; k < PATH_TAIL - 1; k++) for (int j = 0; j < 3; j++) t[k][j] = t[k + 1][j];
      t[PATH_TAIL - 1][0] = x; t[PATH_TAIL - 1][1] = y; t[PATH_TAIL - 1][2] = z;
    } else {
      t[n][0] = x; t[n][1] = y; t[n][2] = z;
    }
    s_lastf[path][0] = (float)x; s_lastf[path][1] = (float)y; s_lastf[path][2] = (float)z;
    if (n < 0x3fffffff) S.n[path] = n + 1;
    double* o = out_paths + ((size_t)path * p.n_frames + f) * 3;
    o[0] = x; o[1] = y; o[2] = z;
    if (out_assign) {
      int8_t* dst = out_assign + ((size_t)path * p.n_frames + f) * C;
      if (C == 8) {  // the 8 nibbles spread to 8 bytes, one store (rows of 8 bytes are 8-byte aligned)
        u64 v = comb & 0xffffffffull;
        v = (v | (v << 16)) & 0x0000ffff0000ffffull;
        v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
        v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
        *reinterpret_cast<u64*>(dst) = v;
      } else {
        for (int c = 0; c < C; c++) dst[c] = (int8_t)((comb >> (4 * c)) & 15);
      }
    }
    if (out_phase) out_phase[(size_t)path * p.n_frames + f] = (uint8_t)phase;
  };
  // The first leaf (priority order) from index `start` on whose mask lies inside the gate g, misses `used`, and whose point is
  // within MAX_STEP of (lx, ly, lz): what the reference's priority_queue pops first that passes :241-246.  One warp, 128
  // leaves per step; the distance runs only for the few leaves inside the gate.
  auto scan = [&](const u64* rec, int L, int start, const u64 (&g)[W], const u64 (&used)[W], double lx, double ly, double lz) {
    int pick = -1;
    for (int i0 = start & ~31; i0 < L && pick < 0; i0 += 128) {
      unsigned cand[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = i0 + 32 * u + lane;
        bool ok = i < L;
        const u64* r = rec + (size_t)(ok ? i : 0) * RW;
#pragma unroll
        for (int w = 0; w < W; w++) { const u64 m = r[w]; ok = ok && !(m & ~g[w]) && !(m & used[w]); }
        cand[u] = __ballot_sync(0xffffffffu, ok);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (cand[u] && pick < 0) {  // cv::norm(c.point - pos) < MAX_STEP, :244
          bool ok = (cand[u] >> lane) & 1u;
          if (ok) {
            const u64* r = rec + (size_t)(i0 + 32 * u + lane) * RW;
            const double x = __longlong_as_double((long long)r[W]) - lx, y = __longlong_as_double((long long)r[W + 1]) - ly,
                         z = __longlong_as_double((long long)r[W + 2]) - lz;
            ok = sqrt_below_step(x * x + y * y + z * z);
          }
          const unsigned hit = __ballot_sync(0xffffffffu, ok);
          if (hit) pick = i0 + 32 * u + __ffs(hit) - 1;
        }
      }
    }
    return pick;
  };

  // frame metadata runs two frames ahead in registers, the staged copy one frame ahead
  const int nf = fb - fa;
  auto meta = [&](int k, int& L, long long& off, int& nd, long long& doff) {
    L = 0; off = 0; nd = 0; doff = 0;
    if (k < nf) { const int g = fa + k - p.f0; L = leaf_cnt[g]; off = leaf_off[g]; nd = fdet_cnt[g]; doff = fdet_off[g]; }
  };
  int L_n1, nd_n1, L_n2, nd_n2;
  long long off_n1, doff_n1, off_n2, doff_n2;
  meta(0, L_n1, off_n1, nd_n1, doff_n1);
  meta(1, L_n2, off_n2, nd_n2, doff_n2);
  if (tid == 0 && nf > 0) stage(0, fa, L_n1, off_n1, nd_n1, doff_n1);

  for (int k = 0; k < nf; k++) {
    const int f = fa + k, b = k & 1;
    const int L = L_n1, nd = nd_n1;
    const long long off = off_n1;
    L_n1 = L_n2; off_n1 = off_n2; nd_n1 = nd_n2; doff_n1 = doff_n2;
    __syncthreads();  // barrier 1 of 2: frame k - 1 is linked (state, its buffer free)
    if (tid == 0 && k + 1 < nf) stage(b ^ 1, f + 1, L_n1, off_n1, nd_n1, doff_n1);
    meta(k + 2, L_n2, off_n2, nd_n2, doff_n2);
    CLS_PROF(0);
    cls_mbar_wait(&full[b], (uint32_t)((k >> 1) & 1));
    CLS_PROF(1);
    const unsigned char* base = link_dyn + (size_t)b * L_::BUF_BYTES;
    const u64* rec = L <= L_::MAX_LEAVES ? reinterpret_cast<const u64*>(base) : leaf_rec + (size_t)off * RW;
    const FrameDet* dets = reinterpret_cast<const FrameDet*>(base + L_::REC_BYTES);
    const int* zs = reinterpret_cast<const int*>(base + L_::REC_BYTES + L_::DET_BYTES);

    // ---- which paths track (:121-123): every warp computes the same list ----
    bool act = false;
    if (lane < D) {
      const int n = S.n[lane];
      const double* last = S.tail[lane][min(max(n, 1), PATH_TAIL) - 1];
      act = n != 0 && !(last[0] == 0 && last[1] == 0 && last[2] == 0);
    }
    const unsigned act_mask = __ballot_sync(0xffffffffu, act);
    const int n_act = __popc(act_mask);
    // this lane's camera (lane < C): the bits of its detections, to count the cameras a gate touches
    u64 cam_bits[W];
    {
      const int b0 = lane < C ? zs[HDR_PREF + lane] : 0, b1 = lane < C ? zs[HDR_PREF + lane + 1] : 0;
#pragma unroll
      for (int w = 0; w < W; w++) {
        const int lo = max(b0 - 64 * w, 0), hi = min(b1 - 64 * w, 64);
        cam_bits[w] = hi > lo ? ((hi == 64 ? ~0ull : ((1ull << hi) - 1)) & ~((1ull << lo) - 1)) : 0ull;
      }
    }
    const int n_slices = (nd + 31) >> 5;

    // ---- speculative phase 1, one warp per tracked path (paths beyond the warp count take turns) ----
    // The ray gate (:228-236) with lane <-> detection: the ballot of one 32-detection test IS a slice of the gate mask.  The
    // distance runs in single precision first and in the reference's double-precision operation order only where single
    // precision cannot decide.  Then the scan for the path's first admissible leaf, IGNORING the picks of earlier paths.
    for (int ai = warp; ai < n_act; ai += LINK_WARPS) {
      int np = 0;
      { unsigned rem = act_mask; for (int q = 0; q < ai; q++) rem &= rem - 1; np = __ffs(rem) - 1; }
      const float lfx = s_lastf[np][0], lfy = s_lastf[np][1], lfz = s_lastf[np][2];
      const double* last = S.tail[np][min(S.n[np], PATH_TAIL) - 1];
      unsigned slice[2 * W];
#pragma unroll
      for (int t = 0; t < 2 * W; t++) {
        slice[t] = 0;
        if (t < n_slices) {
          const int di = 32 * t + lane;
          const bool have = di < nd;
          const FrameDet& fd = dets[have ? di : 0];
          const float wx = lfx - fd.orgf[0], wy = lfy - fd.orgf[1], wz = lfz - fd.orgf[2];
          const float cx = fd.dirf[1] * wz - fd.dirf[2] * wy, cy = fd.dirf[2] * wx - fd.dirf[0] * wz, cz = fd.dirf[0] * wy - fd.dirf[1] * wx;
          const float sf = cx * cx + cy * cy + cz * cz;
          bool gated = sf < (float)(MAX_STEP * MAX_STEP);
          // single precision decides unless sf is within its own error of the threshold: |error of a cross-product component|
          // <= ~6e-7 (|w|_1 |d|_1), and near the threshold d sf = 2 sqrt(sf) d c ~ 700 d c  (taken 3x wider)
          const float tol = 4.f + 1.2e-3f * (fabsf(wx) + fabsf(wy) + fabsf(wz)) * (fabsf(fd.dirf[0]) + fabsf(fd.dirf[1]) + fabsf(fd.dirf[2]));
          if (have && fabsf(sf - (float)(MAX_STEP * MAX_STEP)) < tol) {
            const double ex = last[0] - fd.org[0], ey = last[1] - fd.org[1], ez = last[2] - fd.org[2];  // distToRay, Triangulator.cpp:3-9
            const double fx = fd.dir[1] * ez - fd.dir[2] * ey, fy = fd.dir[2] * ex - fd.dir[0] * ez, fz = fd.dir[0] * ey - fd.dir[1] * ex;
            gated = sqrt_below_step(fx * fx + fy * fy + fz * fz);
          }
          slice[t] = __ballot_sync(0xffffffffu, have && gated);
        }
      }
      u64 g[W], none[W];
      bool touched = false;
#pragma unroll
      for (int w = 0; w < W; w++) { g[w] = ((u64)slice[2 * w + 1] << 32) | slice[2 * w]; none[w] = w == W - 1 ? POISON : 0ull; touched = touched || (g[w] & cam_bits[w]); }
      const int cams_in_gate = __popc(__ballot_sync(0xffffffffu, touched));
      int pick = -1;
      if (cams_in_gate >= MIN_CAMERAS)  // else fillCombinationQueue on the gated container yields nothing
        pick = scan(rec, L, zs[C - cams_in_gate], g, none, last[0], last[1], last[2]);  // leaves with fewer unused cameras cannot lie inside the gate
      if (lane == 0) {
        s_spec[np] = cams_in_gate >= MIN_CAMERAS ? pick : -2;  // -2: nothing gated, no rescan needed
#pragma unroll
        for (int t = 0; t < 2 * W; t++) s_gate[np][t] = slice[t];
      }
    }
    CLS_PROF(2);
    __syncthreads();  // barrier 2 of 2: the speculative picks are in
    if (warp != 0) continue;

    // ---- warp 0: confirm the picks in path order (:119-135) ----
    // A pick that collides with nothing confirmed before it is also the first of the filtered list; otherwise (rare) the
    // path is scanned again with the used mask.  Lane np holds path np's pick.
    u64 used[W];
#pragma unroll
    for (int w = 0; w < W; w++) used[w] = w == W - 1 ? POISON : 0ull;
    unsigned processed = 0;
    {
      int cand = (lane < D && ((act_mask >> lane) & 1u)) ? s_spec[lane] : -2;
      u64 mk[W];
#pragma unroll
      for (int w = 0; w < W; w++) mk[w] = cand >= 0 ? rec[(size_t)cand * RW + w] : 0ull;
      for (unsigned rem = act_mask; rem; rem &= rem - 1) {
        const int np = __ffs(rem) - 1;
        int c_np = __shfl_sync(0xffffffffu, cand, np);
        if (c_np < 0) continue;
        bool clash = false;
#pragma unroll
        for (int w = 0; w < W; w++) clash = clash || (__shfl_sync(0xffffffffu, mk[w], np) & used[w]);
        if (clash) {  // walk the list again with the used filter
          u64 g[W];
#pragma unroll
          for (int w = 0; w < W; w++) g[w] = ((u64)s_gate[np][2 * w + 1] << 32) | s_gate[np][2 * w];
          const double* last = S.tail[np][min(S.n[np], PATH_TAIL) - 1];
          c_np = scan(rec, L, c_np, g, used, last[0], last[1], last[2]);  // nothing before the unfiltered pick can pass
          if (lane == np) {
            cand = c_np;
#pragma unroll
            for (int w = 0; w < W; w++) mk[w] = c_np >= 0 ? rec[(size_t)c_np * RW + w] : 0ull;
          }
          if (c_np < 0) continue;
        }
#pragma unroll
        for (int w = 0; w < W; w++) used[w] |= __shfl_sync(0xffffffffu, mk[w], np);
        processed |= 1u << np;
      }
      if ((processed >> lane) & 1u) emit(lane, f, rec + (size_t)cand * RW, 1);  // all confirmed paths at once, one lane per path
      if (lane == 0) n_phase1 += __popc(processed);
      __syncwarp();
    }
    CLS_PROF(3);
    if (__popc(processed) == D) continue;  // :137

    // ---- phase 2: pickBestCombinations (:200-217), one forward pass with a running used mask, 128 leaves per step ----
    int n_fin = 0;
    for (int i0 = 0; i0 < L; i0 += 128) {
      u64 m[4][W];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = i0 + 32 * u + lane;
#pragma unroll
        for (int w = 0; w < W; w++) m[u][w] = i < L ? rec[(size_t)i * RW + w] : ~0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        bool ok = true;
#pragma unroll
        for (int w = 0; w < W; w++) ok = ok && !(m[u][w] & used[w]);
        unsigned hit = __ballot_sync(0xffffffffu, ok);
        while (hit) {
          const int j = __ffs(hit) - 1;
          bool clash = false;
#pragma unroll
          for (int w = 0; w < W; w++) {
            const u64 mj = __shfl_sync(0xffffffffu, m[u][w], j);
            used[w] |= mj;
            clash = clash || (m[u][w] & mj);
          }
          if (n_fin < LINK_MAX_FINAL) { if (lane == 0) s_fin_idx[n_fin] = i0 + 32 * u + j; n_fin++; }
          else overflow_final = true;
          ok = ok && !clash;  // lane j clashes with itself
          hit = __ballot_sync(0xffffffffu, ok);
        }
      }
    }
    __syncwarp();
    CLS_PROF(4);

    // ---- classifyPaths (:262-332): tail distances one (combination, open path) pair per lane, then per combination its
    // nearest open path (lane <-> combination), the insertion sort and the assignment on lane 0 ----
    unsigned open_paths = 0;  // not processed and not empty: the only ones :269-297 measures
    for (int j = 0; j < D; j++) if (!(processed >> j & 1u) && S.n[j] != 0) open_paths |= 1u << j;
    const int n_open = __popc(open_paths);
    for (int q = lane; q < n_fin * n_open; q += 32) {
      const int ci = q / n_open;
      int j = 0;
      { unsigned rem = open_paths; for (int s2 = q - ci * n_open; s2 > 0; s2--) rem &= rem - 1; j = __ffs(rem) - 1; }
      const int npc = min(S.n[j], PATH_TAIL);
      const u64* r = rec + (size_t)s_fin_idx[ci] * RW;
      const double pt[3] = {__longlong_as_double((long long)r[W]), __longlong_as_double((long long)r[W + 1]), __longlong_as_double((long long)r[W + 2])};
      double dist = 0;
      for (int t = 0; t < npc; t++) dist += dist3(S.tail[j][t], pt);
      s_pdist[ci][j] = dist / (double)npc;
    }
    __syncwarp();
    for (int i = lane; i < n_fin; i += 32) {  // :269-297 (the processed set does not change until the assignment loop)
      int bestPath = 0;
      double bestDist = -1;
      for (unsigned rem = open_paths; rem; rem &= rem - 1) {
        const int j = __ffs(rem) - 1;
        const double dist = s_pdist[i][j];
        if (dist < bestDist || bestDist == -1) { bestDist = dist; bestPath = j; }
      }
      s_cp_path[i] = bestPath; s_cp_err[i] = bestDist;
    }
    __syncwarp();
    // The reference sorts the (combination, nearest path, distance) triples by distance -- std::sort(greater<>) of <= 16
    // elements is libstdc++'s insertion sort: stable, ascending -- and walks them in that order (:299-321).  A triple only
    // acts while an open or an empty path is left, so instead of sorting, the warp extracts the next triple (smallest
    // distance, then smallest index) with shuffles and stops as soon as no path can take a point any more.
    {
      unsigned done = processed;
      unsigned empty_paths = 0;
      for (int i = 0; i < D; i++) if (S.n[i] == 0) empty_paths |= 1u << i;
      unsigned long long taken = 0;  // bit q: triple lane + 32 q of this lane is consumed (n_fin <= 128)
      for (int step = 0; step < n_fin; step++) {
        if (!(open_paths & ~done) && !empty_paths) break;  // every remaining triple would find its path taken and no empty one
        double e = 0;
        int idx = 0x7fffffff;
        for (int q = 0, i = lane; i < n_fin; q++, i += 32)
          if (!(taken >> q & 1ull) && (idx == 0x7fffffff || s_cp_err[i] < e)) { e = s_cp_err[i]; idx = i; }
        for (int o = 16; o > 0; o >>= 1) {
          const double e2 = __shfl_xor_sync(0xffffffffu, e, o);
          const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
          if (i2 != 0x7fffffff && (idx == 0x7fffffff || e2 < e || (e2 == e && i2 < idx))) { e = e2; idx = i2; }
        }
        if ((idx & 31) == lane) taken |= 1ull << (idx >> 5);
        const int pth = s_cp_path[idx];
        int target = -1;
        if (done >> pth & 1u) { if (empty_paths) target = __ffs(empty_paths) - 1; }  // the first empty path (:305-311)
        else target = pth;
        if (target != -1) {
          if (lane == 0) { emit(target, f, rec + (size_t)s_fin_idx[idx] * RW, 2); n_phase2++; }
          done |= 1u << target;
          empty_paths &= ~(1u << target);
          __syncwarp();
        }
      }
    }
    __syncwarp();
    CLS_PROF(5);
  }
  __syncthreads();
  for (int i = tid; i < (int)(sizeof(LinkState) / sizeof(int)); i += LINK_THREADS) ((int*)st)[i] = ((const int*)&S)[i];
#ifdef TRI_TUNING
  if (tid == 0) for (int q = 0; q < 8; q++) atomicAdd(&ctr->prof[q], prof_acc[q]);
#endif
  if (tid == 0) {
    atomicAdd(&ctr->phase1, n_phase1); atomicAdd(&ctr->phase2, n_phase2);
    if (overflow_final) atomicExch(&ctr->overflow_final, 1);
  }
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
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  double enumerate_ms = 0, link_ms = 0;  // of the call in progress
  ~ClsWork() { for (cudaEvent_t v : ev) if (v) cudaEventDestroy(v); }
  cudaError_t events() {
    for (cudaEvent_t& v : ev)
      if (!v) { cudaError_t err = cudaEventCreate(&v); if (err != cudaSuccess) return err; }
    return cudaSuccess;
  }
  DevBuf offs, dets, paths, assign, phase, state, ctr, front, txyz, terr, lrec, loff, lcnt, hdr, fdet, fdoff, fdcnt, seq;
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

namespace tri {
cudaError_t launch_lazy_link(cudaStream_t s, const DltRig<double>& dlt, const RayRig& ray, const ClsParams& p, int n_seq, const int2* d_seq,
                             const int32_t* d_offs, const double* d_dets, LinkState* d_state, double* d_paths, int8_t* d_assign,
                             uint8_t* d_phase, ClsCounters* d_ctr);
}

namespace {

// argument checks shared by the entry points; *n_det = detections in the CSR
int cls_check(tri_engine* e, int mode, int n_drones, const int32_t* det_offsets, const double* dets_xy, int n_frames, int64_t* n_det,
              bool lazy = false) {
  if (!e) return fail(TRI_ERR_ARG, "null engine");
  if (mode != TRI_MATRIX && mode != TRI_RAY) return fail(TRI_ERR_ARG, "mode must be TRI_MATRIX or TRI_RAY");
  const int C = e->n_cams;
  if (C > CLS_MAX_CAMS && !lazy)
    return fail(TRI_ERR_ARG, "the frame-sharded classifier enumerates the candidate combinations, which is exponential in the camera count: at most "
                             "16 cameras (tri_classify / tri_classify_sequences take up to 32 through the lazy search)");
  if (n_drones < 1 || n_drones > TRI_MAX_DRONES) return fail(TRI_ERR_ARG, "n_drones must be in [1, TRI_MAX_DRONES]");
  if (n_frames < 0) return fail(TRI_ERR_ARG, "bad frame count");
  *n_det = 0;
  if (n_frames == 0) return TRI_OK;
  if (!det_offsets) return fail(TRI_ERR_ARG, "null detection offsets");
  for (int c = 0; c < C; c++) {
    const int32_t* o = det_offsets + (size_t)c * (n_frames + 1);
    for (int f = 0; f < n_frames; f++) {
      if (o[f + 1] < o[f]) return fail(TRI_ERR_ARG, "detection offsets must be non-decreasing");
      if (o[f + 1] - o[f] > TRI_MAX_DETS) return fail(TRI_ERR_CAPACITY, "more than TRI_MAX_DETS detections on one camera in one frame");
    }
    *n_det = std::max<int64_t>(*n_det, o[n_frames]);
  }
  if (*n_det > 0 && !dets_xy) return fail(TRI_ERR_ARG, "null detections");
  return TRI_OK;
}

ClsParams cls_params(const tri_engine* e, int mode, unsigned flags, int n_drones, int n_frames) {
  ClsParams p{};
  p.n_cams = e->n_cams; p.n_drones = n_drones; p.n_frames = n_frames;
  p.solver = mode == TRI_MATRIX ? 0 : (flags & TRI_RAY_CLOSED_FORM) ? 2 : 1;  // ray: the reference's LM trajectory unless the fast solver is asked for
  p.error_ = mode == TRI_MATRIX ? MAX_ERROR_MATRIX : MAX_ERROR_RAY;  // DroneClassifier.cpp:3-10
  p.W = e->n_cams <= 8 ? 2 : 4;
  return p;
}

int cls_grid(const tri_engine* e, int frames) {
  // every resident CTA slot gets a frame: the tree expansion is latency-bound (dependent FP64 chains, idle
  // lanes on narrow levels), so occupancy is what hides it -- 4 CTAs per SM instead of 2: S09_D6 with the
  // exact LM 3.88 -> 2.82 s (profiles/r1_cls_grid_sweep.log)
  int per_sm = 2;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, enumerate_kernel, CLS_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
#ifdef TRI_TUNING
  if (const char* v = getenv("TRI_CLS_CTAS_PER_SM")) per_sm = std::max(1, atoi(v));
#endif
  return std::max(1, std::min(frames, per_sm * e->sm_count));
}

// (A) on frames [p.f0, p.f1) with the work buffers sized by (cap, leaf_cap); *h = the counters after the launch
int cls_enumerate(tri_engine* e, ClsWork& W, ClsParams& p, int cap, long long leaf_cap, int64_t n_det_batch, ClsCounters* h) {
  cudaStream_t s = e->stream;
  const int frames = p.f1 - p.f0, grid = cls_grid(e, frames);
  TRI_CUDA(W.front.alloc(sizeof(u64) * 2 * (size_t)cap * grid));
  TRI_CUDA(W.txyz.alloc(sizeof(double) * 3 * (size_t)cap * grid));
  TRI_CUDA(W.terr.alloc(sizeof(double) * (size_t)cap * grid));
  TRI_CUDA(W.lrec.alloc(sizeof(u64) * rec_words(p.W) * (size_t)leaf_cap));
  TRI_CUDA(W.loff.alloc(sizeof(long long) * frames));
  TRI_CUDA(W.lcnt.alloc(sizeof(int) * frames));
  TRI_CUDA(W.hdr.alloc(sizeof(int) * HDR_INTS * (size_t)frames));
  TRI_CUDA(W.fdet.alloc(sizeof(FrameDet) * (size_t)std::max<int64_t>(n_det_batch, 1)));
  TRI_CUDA(W.fdoff.alloc(sizeof(long long) * frames));
  TRI_CUDA(W.fdcnt.alloc(sizeof(int) * frames));
  p.cap = cap; p.leaf_cap = leaf_cap;
  ClsCounters* ctr = W.ctr.as<ClsCounters>();
  TRI_CUDA(cudaMemsetAsync(&ctr->leaf_total, 0, 2 * sizeof(u64), s));  // leaf_total, fdet_total: offsets within this batch
  TRI_CUDA(W.events());
  TRI_CUDA(cudaEventRecord(W.ev[0], s));
  enumerate_kernel<<<grid, CLS_THREADS, 0, s>>>(e->rig64, e->ray, p, W.offs.as<int32_t>(), W.dets.as<double>(), W.front.as<u64>(),
                                                 W.txyz.as<double>(), W.terr.as<double>(), W.lrec.as<u64>(), W.loff.as<long long>(),
                                                 W.lcnt.as<int>(), W.hdr.as<int>(), W.fdet.as<FrameDet>(), W.fdoff.as<long long>(),
                                                 W.fdcnt.as<int>(), ctr);
  e->launches++;
  TRI_CUDA(cudaGetLastError());
  TRI_CUDA(cudaEventRecord(W.ev[1], s));
  TRI_CUDA(cudaMemcpyAsync(h, ctr, sizeof(*h), cudaMemcpyDeviceToHost, s));
  TRI_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  TRI_CUDA(cudaEventElapsedTime(&ms, W.ev[0], W.ev[1]));
  W.enumerate_ms += ms;
  return TRI_OK;
}

// (B) on the batch last enumerated: one warp per sequence (seq == nullptr: the single sequence [p.f0, p.f1))
int cls_link(tri_engine* e, ClsWork& W, const ClsParams& p, int n_seq, const int2* d_seq, int first_seq, bool want_assign, bool want_phase) {
  cudaStream_t s = e->stream;
  auto go = [&](auto kern, int bytes) -> int {
    TRI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    TRI_CUDA(W.events());
    TRI_CUDA(cudaEventRecord(W.ev[1], s));
    kern<<<n_seq, LINK_THREADS, bytes, s>>>(e->ray, p, d_seq, W.lrec.as<u64>(), W.loff.as<long long>(), W.lcnt.as<int>(), W.hdr.as<int>(),
                                  W.fdet.as<FrameDet>(), W.fdoff.as<long long>(), W.fdcnt.as<int>(), W.state.as<LinkState>() + first_seq,
                                  W.paths.as<double>(), want_assign ? W.assign.as<int8_t>() : nullptr,
                                  want_phase ? W.phase.as<uint8_t>() : nullptr, W.ctr.as<ClsCounters>());
    e->launches++;
    TRI_CUDA(cudaGetLastError());
    TRI_CUDA(cudaEventRecord(W.ev[2], s));
    TRI_CUDA(cudaEventSynchronize(W.ev[2]));  // the next batch's enumeration reuses the buffers this pass reads
    float ms = 0;
    TRI_CUDA(cudaEventElapsedTime(&ms, W.ev[1], W.ev[2]));
    W.link_ms += ms;
    return TRI_OK;
  };
  return p.W == 2 ? go(link_kernel<2>, LinkLayout<2>::BYTES) : go(link_kernel<4>, LinkLayout<4>::BYTES);
}

int64_t dets_in_frames(const int32_t* det_offsets, int C, int n_frames, int f0, int f1) {
  int64_t n = 0;
  for (int c = 0; c < C; c++) { const int32_t* o = det_offsets + (size_t)c * (n_frames + 1); n += o[f1] - o[f0]; }
  return n;
}

void cls_stats(tri_classify_stats* stats, const ClsCounters& h, int max_frontier, const ClsWork& W) {
  if (!stats) return;
  stats->enumerate_us = (int64_t)(W.enumerate_ms * 1e3); stats->link_us = (int64_t)(W.link_ms * 1e3);
  stats->nodes = (int64_t)h.nodes; stats->solves = (int64_t)h.solves; stats->leaves = (int64_t)h.leaves;
  stats->lm_iters = (int64_t)h.lm_iters; stats->phase1 = (int64_t)h.phase1; stats->phase2 = (int64_t)h.phase2;
  stats->ties = (int64_t)h.ties; stats->max_frontier = max_frontier;
}

// Classify n_seq independent sequences laid back to back in one CSR of n_frames frames (seq_bounds[s] .. seq_bounds[s+1]).
int cls_run(tri_engine* e, int mode, unsigned flags, int n_drones, const int32_t* det_offsets, const double* dets_xy, int n_frames,
            int n_seq, const int32_t* seq_bounds, double* out_paths, int8_t* out_assign, uint8_t* out_phase, tri_classify_stats* stats) {
  int64_t n_det = 0;
  const bool lazy = e && (e->n_cams > CLS_MAX_CAMS || (flags & TRI_CLS_LAZY));
  int st = cls_check(e, mode, n_drones, det_offsets, dets_xy, n_frames, &n_det, lazy);
  if (st != TRI_OK) return st;
  if (!out_paths && n_frames > 0) return fail(TRI_ERR_ARG, "bad output arguments");
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n_frames == 0) return TRI_OK;
  const int C = e->n_cams;
  DeviceGuard g(e->device);
  cudaStream_t s = e->stream;
  ClsParams p = cls_params(e, mode, flags, n_drones, n_frames);
  if (!e->cls_work) { e->cls_work = new ClsWork(); e->cls_work_free = free_cls_work; }
  ClsWork& W = *static_cast<ClsWork*>(e->cls_work);
  W.pending = false;
  W.enumerate_ms = W.link_ms = 0;
  const size_t n_offs = (size_t)C * (n_frames + 1);
  const size_t sz_paths = sizeof(double) * 3 * n_drones * (size_t)n_frames, sz_assign = (size_t)n_drones * n_frames * C,
               sz_phase = (size_t)n_drones * n_frames;
  const bool multi = n_seq > 1;
  TRI_CUDA(W.offs.alloc(sizeof(int32_t) * n_offs));
  TRI_CUDA(W.dets.alloc(sizeof(double) * 2 * n_det));
  TRI_CUDA(W.paths.alloc(sz_paths));
  TRI_CUDA(W.assign.alloc(sz_assign));
  TRI_CUDA(W.phase.alloc(sz_phase));
  TRI_CUDA(W.state.alloc(sizeof(LinkState) * (size_t)std::max(n_seq, 1)));
  TRI_CUDA(W.ctr.alloc(sizeof(ClsCounters)));
  TRI_CUDA(cudaMemcpyAsync(W.offs.p, det_offsets, sizeof(int32_t) * n_offs, cudaMemcpyHostToDevice, s));
  if (n_det) TRI_CUDA(cudaMemcpyAsync(W.dets.p, dets_xy, sizeof(double) * 2 * n_det, cudaMemcpyHostToDevice, s));
  TRI_CUDA(cudaMemsetAsync(W.paths.p, 0, sz_paths, s));
  TRI_CUDA(cudaMemsetAsync(W.assign.p, 0xff, sz_assign, s));
  TRI_CUDA(cudaMemsetAsync(W.phase.p, 0, sz_phase, s));
  TRI_CUDA(cudaMemsetAsync(W.state.p, 0, sizeof(LinkState) * (size_t)std::max(n_seq, 1), s));
  TRI_CUDA(cudaMemsetAsync(W.ctr.p, 0, sizeof(ClsCounters), s));
  std::vector<int2> seq;
  if (multi) {  // every sequence links on its own warp: all frames are enumerated first, in one batch
    for (int q = 0; q < n_seq; q++) {
      if (seq_bounds[q] < 0 || seq_bounds[q + 1] < seq_bounds[q] || seq_bounds[q + 1] > n_frames) return fail(TRI_ERR_ARG, "sequence bounds must be non-decreasing within [0, n_frames]");
      seq.push_back(make_int2(seq_bounds[q], seq_bounds[q + 1]));
    }
    TRI_CUDA(W.seq.alloc(sizeof(int2) * n_seq));
    TRI_CUDA(cudaMemcpyAsync(W.seq.p, seq.data(), sizeof(int2) * n_seq, cudaMemcpyHostToDevice, s));
  }

  if (lazy) {  // more than 16 cameras, or on request: no enumeration, the lazy best-first search of tri_classify_lazy.cu
    p.f0 = 0; p.f1 = n_frames;
    TRI_CUDA(W.events());
    TRI_CUDA(cudaEventRecord(W.ev[1], s));
    TRI_CUDA(launch_lazy_link(s, e->rig64, e->ray, p, multi ? n_seq : 1, multi ? W.seq.as<int2>() : nullptr, W.offs.as<int32_t>(), W.dets.as<double>(),
                              W.state.as<LinkState>(), W.paths.as<double>(), out_assign ? W.assign.as<int8_t>() : nullptr,
                              out_phase ? W.phase.as<uint8_t>() : nullptr, W.ctr.as<ClsCounters>()));
    e->launches++;
    TRI_CUDA(cudaEventRecord(W.ev[2], s));
    TRI_CUDA(cudaMemcpyAsync(out_paths, W.paths.p, sz_paths, cudaMemcpyDeviceToHost, s));
    if (out_assign) TRI_CUDA(cudaMemcpyAsync(out_assign, W.assign.p, sz_assign, cudaMemcpyDeviceToHost, s));
    if (out_phase) TRI_CUDA(cudaMemcpyAsync(out_phase, W.phase.p, sz_phase, cudaMemcpyDeviceToHost, s));
    ClsCounters hz{};
    TRI_CUDA(cudaMemcpyAsync(&hz, W.ctr.p, sizeof(hz), cudaMemcpyDeviceToHost, s));
    TRI_CUDA(cudaStreamSynchronize(s));
    float ms = 0;
    TRI_CUDA(cudaEventElapsedTime(&ms, W.ev[1], W.ev[2]));
    W.link_ms = ms;
    if (hz.bad_input) return fail(TRI_ERR_ARG, "malformed detection offsets");
    if (hz.overflow_frontier) return fail(TRI_ERR_CAPACITY, "the lazy search visited more than 2^20 nodes for one combination");
    if (hz.overflow_final) return fail(TRI_ERR_CAPACITY, "more than 128 disjoint combinations kept in one frame");
    cls_stats(stats, hz, 0, W);
    return TRI_OK;
  }

  // Work is cut into batches of frames: (A) enumerates a batch, (B) links it.  One sequence: 8192 frames per batch, the
  // state carried from batch to batch.  Many sequences: whole sequences per batch (about 128 k frames), one CTA each.
  int batch = std::min(n_frames, 8192);
  int cap = 1 << 14;
  long long leaf_cap = 4ll << 20;
  ClsCounters h{};
  int max_frontier = 0;
  int q0 = 0;  // first sequence of the batch (multi)
  for (int f0 = 0; f0 < n_frames;) {
    int f1 = std::min(n_frames, f0 + batch), q1 = q0;
    if (multi) {
      q1 = q0 + 1;
      while (q1 < n_seq && seq_bounds[q1 + 1] - seq_bounds[q0] <= (128 << 10)) q1++;
      f1 = seq_bounds[q1];
      if (f1 == f0) { q0 = q1; continue; }  // empty sequences
      leaf_cap = std::max<long long>(leaf_cap, 1024ll * (f1 - f0));
    }
    p.f0 = f0; p.f1 = f1;
    ClsCounters before;
    TRI_CUDA(cudaMemcpyAsync(&before, W.ctr.p, sizeof(before), cudaMemcpyDeviceToHost, s));
    TRI_CUDA(cudaStreamSynchronize(s));
    if ((st = cls_enumerate(e, W, p, cap, leaf_cap, dets_in_frames(det_offsets, C, n_frames, f0, f1), &h)) != TRI_OK) return st;
    if (h.bad_input) return fail(TRI_ERR_ARG, "malformed detection offsets");
    if (h.overflow_frontier || h.overflow_leaves) {
      // grow the work buffers (or shrink the batch) and redo this batch; the statistics of the aborted attempt are rolled back
      if (h.overflow_frontier) { if (cap >= (1 << 22)) return fail(TRI_ERR_CAPACITY, "combination frontier exceeds 4M nodes in one frame"); cap *= 4; if (!multi) batch = std::max(1, batch / 4); }
      if (h.overflow_leaves) { if (!multi && batch > 64) batch /= 4; else if (leaf_cap < (1ll << 30)) leaf_cap *= 4; else return fail(TRI_ERR_CAPACITY, "candidate list exceeds device buffer"); }
      before.overflow_frontier = before.overflow_leaves = 0;
      TRI_CUDA(cudaMemcpyAsync(W.ctr.p, &before, sizeof(before), cudaMemcpyHostToDevice, s));
      TRI_CUDA(cudaStreamSynchronize(s));
      continue;
    }
    max_frontier = std::max(max_frontier, h.max_frontier);
    if ((st = cls_link(e, W, p, multi ? q1 - q0 : 1, multi ? W.seq.as<int2>() + q0 : nullptr, multi ? q0 : 0, out_assign != nullptr, out_phase != nullptr)) != TRI_OK) return st;
    f0 = f1;
    q0 = q1;
  }
  TRI_CUDA(cudaMemcpyAsync(out_paths, W.paths.p, sz_paths, cudaMemcpyDeviceToHost, s));
  if (out_assign) TRI_CUDA(cudaMemcpyAsync(out_assign, W.assign.p, sz_assign, cudaMemcpyDeviceToHost, s));
  if (out_phase) TRI_CUDA(cudaMemcpyAsync(out_phase, W.phase.p, sz_phase, cudaMemcpyDeviceToHost, s));
  TRI_CUDA(cudaMemcpyAsync(&h, W.ctr.p, sizeof(h), cudaMemcpyDeviceToHost, s));
  TRI_CUDA(cudaStreamSynchronize(s));
  if (h.overflow_final) return fail(TRI_ERR_CAPACITY, "more than 128 disjoint combinations kept in one frame");
#ifdef TRI_TUNING
  if (getenv("TRI_CLS_PROFILE"))
    fprintf(stderr, "link cycles: top %llu | wait %llu | gates %llu | phase1 %llu | phase2 %llu | classifyPaths %llu\n", h.prof[0], h.prof[1],
            h.prof[2], h.prof[3], h.prof[4], h.prof[5]);
#endif
  cls_stats(stats, h, max_frontier, W);
  return TRI_OK;
}

}  // namespace

extern "C" int tri_classify(tri_engine* e, int mode, unsigned flags, int n_drones, const int32_t* det_offsets,
                            const double* dets_xy, int n_frames, double* out_paths, int8_t* out_assign, uint8_t* out_phase,
                            tri_classify_stats* stats) {
  return cls_run(e, mode, flags, n_drones, det_offsets, dets_xy, n_frames, 1, nullptr, out_paths, out_assign, out_phase, stats);
}

// Many independent sequences (recordings) in one call: sequence q holds the frames [seq_bounds[q], seq_bounds[q+1]) of the
// CSR; every frame of every sequence is enumerated in one launch and each sequence is linked by its own warp.
extern "C" int tri_classify_sequences(tri_engine* e, int mode, unsigned flags, int n_drones, int n_seq, const int32_t* seq_bounds,
                                      const int32_t* det_offsets, const double* dets_xy, int n_frames, double* out_paths,
                                      int8_t* out_assign, uint8_t* out_phase, tri_classify_stats* stats) {
  if (n_seq < 1 || !seq_bounds) return fail(TRI_ERR_ARG, "no sequences");
  if (seq_bounds[0] != 0 || seq_bounds[n_seq] != n_frames) return fail(TRI_ERR_ARG, "sequence bounds must cover [0, n_frames]");
  if (n_seq == 1) return cls_run(e, mode, flags, n_drones, det_offsets, dets_xy, n_frames, 1, nullptr, out_paths, out_assign, out_phase, stats);
  return cls_run(e, mode, flags, n_drones, det_offsets, dets_xy, n_frames, n_seq, seq_bounds, out_paths, out_assign, out_phase, stats);
}

// ---- frame-sharded classification: candidate generation is independent per frame, linking is a
// chain.  tri_classify_begin enumerates a contiguous shard of the sequence and keeps every frame's candidate
// list on the device; tri_classify_finish links the shard starting from the tracking state the previous shard
// ended with (n_drones x (last point + 3-point tail) + counters, an opaque tri_classify_state_bytes() blob) and
// returns the state for the next one.  All ranks enumerate at once; only the small state travels along the chain.
extern "C" int tri_classify_state_bytes(void) { return (int)sizeof(LinkState); }

extern "C" int tri_classify_begin(tri_engine* e, int mode, unsigned flags, int n_drones, const int32_t* det_offsets,
                                  const double* dets_xy, int n_frames) {
  int64_t n_det = 0;
  int st = cls_check(e, mode, n_drones, det_offsets, dets_xy, n_frames, &n_det);
  if (st != TRI_OK) return st;
  const int C = e->n_cams;
  DeviceGuard g(e->device);
  cudaStream_t s = e->stream;
  if (!e->cls_work) { e->cls_work = new ClsWork(); e->cls_work_free = free_cls_work; }
  ClsWork& W = *static_cast<ClsWork*>(e->cls_work);
  W.pending = false;
  W.enumerate_ms = W.link_ms = 0;
  ClsParams p = cls_params(e, mode, flags, n_drones, n_frames);
  p.f0 = 0; p.f1 = n_frames;
  W.job = p;
  W.job_max_frontier = 0;
  if (n_frames == 0) { W.pending = true; return TRI_OK; }
  const size_t n_offs = (size_t)C * (n_frames + 1);
  TRI_CUDA(W.offs.alloc(sizeof(int32_t) * n_offs));
  TRI_CUDA(W.dets.alloc(sizeof(double) * 2 * n_det));
  TRI_CUDA(W.ctr.alloc(sizeof(ClsCounters)));
  TRI_CUDA(cudaMemcpyAsync(W.offs.p, det_offsets, sizeof(int32_t) * n_offs, cudaMemcpyHostToDevice, s));
  if (n_det) TRI_CUDA(cudaMemcpyAsync(W.dets.p, dets_xy, sizeof(double) * 2 * n_det, cudaMemcpyHostToDevice, s));
  int cap = 1 << 14;
  long long leaf_cap = std::max<long long>(4ll << 20, 1024ll * n_frames);
  for (;;) {  // the whole shard's candidates stay resident: on overflow grow the buffers and enumerate again
    TRI_CUDA(cudaMemsetAsync(W.ctr.p, 0, sizeof(ClsCounters), s));
    ClsCounters h{};
    if ((st = cls_enumerate(e, W, p, cap, leaf_cap, n_det, &h)) != TRI_OK) return st;
    if (h.bad_input) return fail(TRI_ERR_ARG, "malformed detection offsets");
    if (h.overflow_frontier) { if (cap >= (1 << 22)) return fail(TRI_ERR_CAPACITY, "combination frontier exceeds 4M nodes in one frame"); cap *= 4; continue; }
    if (h.overflow_leaves) { if (leaf_cap >= (1ll << 30)) return fail(TRI_ERR_CAPACITY, "candidate list exceeds device buffer"); leaf_cap *= 4; continue; }
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
  int st = cls_link(e, W, p, 1, nullptr, 0, out_assign != nullptr, out_phase != nullptr);
  if (st != TRI_OK) return st;
  ClsCounters h{};
  TRI_CUDA(cudaMemcpyAsync(out_paths, W.paths.p, sz_paths, cudaMemcpyDeviceToHost, s));
  if (out_assign) TRI_CUDA(cudaMemcpyAsync(out_assign, W.assign.p, sz_assign, cudaMemcpyDeviceToHost, s));
  if (out_phase) TRI_CUDA(cudaMemcpyAsync(out_phase, W.phase.p, sz_phase, cudaMemcpyDeviceToHost, s));
  if (state_out) TRI_CUDA(cudaMemcpyAsync(state_out, W.state.p, sizeof(LinkState), cudaMemcpyDeviceToHost, s));
  TRI_CUDA(cudaMemcpyAsync(&h, W.ctr.p, sizeof(h), cudaMemcpyDeviceToHost, s));
  TRI_CUDA(cudaStreamSynchronize(s));
  if (h.overflow_final) return fail(TRI_ERR_CAPACITY, "more than 128 disjoint combinations kept in one frame");
  cls_stats(stats, h, W.job_max_frontier, W);
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
      stats->enumerate_us = std::max(stats->enumerate_us, st.enumerate_us); stats->link_us += st.link_us;
      stats->max_frontier = std::max(stats->max_frontier, st.max_frontier);
    }
  }
  return TRI_OK;
}
