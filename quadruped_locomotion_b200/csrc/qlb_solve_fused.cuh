// qlb_solve_fused.cuh - ONE persistent kernel for the whole contact-force pipeline of a batch.
//
// Every warp is autonomous (no CTA-wide synchronisation after the prologue) and alternates between two
// phases:
//
//   tile phase   eight consecutive states (one leg per lane, as in qlb_solve_quad.cuh): the raw SoA input
//                rows of the tile have been staged in shared memory by the TMA unit (one
//                cp.async.bulk.tensor box {8 states x rows} per input array, completion on a per-warp mbarrier)
//                while the previous tile was computed.  Kinematics, friction frames, wrench map, the
//                unconstrained minimiser.  States whose minimiser is feasible are finished.  The others
//                ("hard" states, 28 % of config C3) are parked in the warp's shared-memory STASH together with
//                everything the rounds need (friction frame, foot position, Jacobian, gravity torques, ...):
//                nothing is recomputed and nothing goes through HBM.
//   round phase  entered when the stash holds a warp's worth of hard states: each quad takes one state
//                and runs equality-constrained rounds on the 6x6 dual system.  A quad that finishes writes its
//                outputs and takes the next state from the stash, so the eight quads of the warp stay busy;
//                when the stash runs low the unfinished states are written back (iterate, multipliers,
//                pattern) and the warp returns to the tile phase.
//
// The rounds are a DUAL BLOCK ACTIVE-SET method (the Goldfarb-Idnani idea with block additions; the dual of
// the QP is a bound-constrained QP in the multipliers and this is the primal active-set method on it):
//   state  (y, u, F): primal iterate, multipliers u >= 0 of the rows in the working set F
//   round  (y+, u+) = solution of the equality-constrained QP for F (the 6x6 Woodbury system);
//          ratio test: the largest step t in [0, 1] towards (y+, u+) that keeps every multiplier >= 0;
//          t < 1: move by t, drop the blocking rows from F;
//          t = 1: move; every violated row joins F (multiplier 0); no violated row -> optimal.
// Each non-zero step increases the dual objective, so no working set repeats: finite, no cycling, and in
// the common cases it takes the same steps as the "repair every violated row" heuristic it replaces (which
// needed an interior-point pass as a safety net for 1 % of the states).  Measured on 60 000 C3 / C5 states
// (tools/proto/rounds_proto.py): 1.69 / 1.45 rounds per hard state, at most 14, no failure, active set equal to
// the reference solver's on every state.  States that exhaust the round limit (none seen) go to a.list2 for
// the interior-point kernel of qlb_solve_quad.cuh.
//
// Reference path: ContactForceDistribution.cpp:99-136,138-336,385-578,614-625;
// quadrupedkinematics.cpp:143-278,485-552; VirtualModelController.cpp:89-268 (as in qlb_solve_quad.cuh).
#pragma once

#include <cuda.h>

#include "qlb_solve_quad.cuh"

namespace qlb {

#ifndef QLB_FUSED_MIN_CTAS
#define QLB_FUSED_MIN_CTAS 3
#endif
#ifndef QLB_DBAS_MAX_ROUNDS
#define QLB_DBAS_MAX_ROUNDS 40
#endif
constexpr int kFusedSmemBudget = 75 * 1024;   // per CTA: three CTAs per SM (228 KB, 1 KB reserved per CTA)

// ---------------------------------------------------------------------------------------------------------
// Staging of the raw input rows: one segment per input array, each a dense [rows][8 states] box.
template <typename real, int MODE>
struct Staging {
  static constexpr int kRow = 8 * (int)sizeof(real);                           // bytes of one row of a box
  __host__ __device__ static constexpr int align(int x) { return (x + 127) & ~127; }              // TMA destinations: 128-byte aligned
  // wrench mode: q, quat, wrench, mu, normals;  state mode: q, pose, twist, tpose, ttwist, mu, normals
  static constexpr int kNumSeg = (MODE == 1) ? 7 : 5;
  __host__ __device__ static constexpr int rows(int s) {
    return (MODE == 1) ? (s == 0 ? 12 : s == 1 ? 7 : s == 2 ? 6 : s == 3 ? 7 : s == 4 ? 6 : s == 5 ? 4 : 12)
                       : (s == 0 ? 12 : s == 1 ? 4 : s == 2 ? 6 : s == 3 ? 4 : 12);
  }
  __host__ __device__ static constexpr int offset(int s) { return s == 0 ? 0 : offset(s - 1) + align(rows(s - 1) * kRow); }
  static constexpr int kBytes = offset(kNumSeg - 1) + align(rows(kNumSeg - 1) * kRow);
  static constexpr int kSegMu = (MODE == 1) ? 5 : 3;
  static constexpr int kSegNormals = (MODE == 1) ? 6 : 4;
};

// Tensor maps of the input arrays of one call (built on the host, qlb_api.cu), in Staging segment order.
struct alignas(64) FusedMaps {
  CUtensorMap seg[7];
};

// ---------------------------------------------------------------------------------------------------------
// The stash of one warp: CAP entries.  Per-lane planes (element k of the entry in slot s, leg l at
// plane[k * 4 CAP + 4 s + l]) and a per-entry header.
constexpr int kStashLane = 19;   // friction frame / force rows of the wrench map (9), foot (3), mu, y (3), u (3)
template <typename real, typename creal, int CAP>
struct StashLayout {
  static constexpr int kQ = 4 * CAP;
  static constexpr int kLaneBytes = kStashLane * kQ * (int)sizeof(creal);
  static constexpr int kJBytes = 12 * kQ * (int)sizeof(real);
  static constexpr int kBBytes = 6 * CAP * (int)sizeof(creal);
  static constexpr int kHBytes = 4 * CAP * 4;
  static constexpr int kBytes = ((kLaneBytes + kJBytes + kBBytes + kHBytes) + 15) & ~15;
};

template <typename real, typename creal, int MODE>
struct FusedLayout {
  // fixed part: parameter block, solver constants, one mbarrier per warp (the leg-model table is a static array)
  static constexpr int kFixed = ((((int)sizeof(DeviceParamsT<real>) + 15) & ~15) + (((int)sizeof(CoreConst<creal>) + 15) & ~15) + 64 + 127) & ~127;
  static constexpr int kStatic = (int)sizeof(DeviceModelT<double>) + 128;
  static constexpr int kStage = Staging<real, MODE>::kBytes;
  static constexpr int stash_bytes(int c) {
    return ((kStashLane * 4 * c * (int)sizeof(creal) + 12 * 4 * c * (int)sizeof(real) + 6 * c * (int)sizeof(creal) + 16 * c) + 15) & ~15;
  }
  static constexpr int warp_bytes(int c) { return (kStage + stash_bytes(c) + 127) & ~127; }
  static constexpr int cap_for(int c) { return (kStatic + kFixed + 4 * warp_bytes(c) <= kFusedSmemBudget || c <= 8) ? c : cap_for(c - 1); }
  static constexpr int kCap = cap_for(15);
  static_assert(kCap >= 10, "stash too small");
  static_assert(stash_bytes(kCap) == StashLayout<real, creal, kCap>::kBytes, "layout mismatch");
  static constexpr int kWarpBytes = warp_bytes(kCap);
  static constexpr int kTotal = kFixed + 4 * kWarpBytes;   // dynamic shared memory of the kernel
  static_assert(kStatic + kTotal <= kFusedSmemBudget, "shared memory budget");
};

template <typename real, typename creal, int CAP>
struct WarpStash {
  creal* sl;      // [kStashLane][4 CAP]
  real* sj;       // [12][4 CAP]   Jacobian (9) and gravity torques (3); also the scratch of the tile phase
  creal* sb;      // [6][CAP]      desired wrench
  uint32_t* sh;   // [4][CAP]      state index; mask | pattern << 4 | rounds << 24; gradient scale (float bits); spare
  __device__ WarpStash(unsigned char* base) {
    using SL = StashLayout<real, creal, CAP>;
    sl = reinterpret_cast<creal*>(base);
    sj = reinterpret_cast<real*>(base + SL::kLaneBytes);
    sb = reinterpret_cast<creal*>(base + SL::kLaneBytes + SL::kJBytes);
    sh = reinterpret_cast<uint32_t*>(base + SL::kLaneBytes + SL::kJBytes + SL::kBBytes);
  }
};

// ---------------------------------------------------------------------------------------------------------
// mbarrier / TMA / cp.async primitives (shared::cta addresses as 32-bit)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0u;
}
// Bounded: a transfer that never completes (a bad tensor map) ends the kernel with an error instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_elem(uint32_t dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_elem(uint32_t dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// the j-th (0-based) set bit of m, or -1
__device__ __forceinline__ int nth_set_bit(unsigned m, const int j) {
#pragma unroll
  for (int i = 0; i < 7; i++)
    if (i < j) m &= m - 1u;
  return m ? (__ffs(m) - 1) : -1;
}

// Issue the loads of tile `tile` (eight states) into this warp's staging buffer.
// TMA: lane 0 arms the mbarrier with the byte count and issues one box per input array (out-of-range
// columns of the last tile are zero-filled by the unit).  Otherwise (batch size or pointers not 16-byte
// aligned): every lane copies its share with 8-/4-byte cp.async, column index clamped to B - 1.
template <typename real, int MODE, bool TMA>
__device__ __forceinline__ void stage_issue(const SolveArgsT<real>& a, const FusedMaps& maps, const unsigned long long tile,
                                            unsigned char* stage, const uint32_t bar, const int lane) {
  using SG = Staging<real, MODE>;
  const real* src[7];
  if (MODE == 1) { src[0] = a.q; src[1] = a.pose; src[2] = a.twist; src[3] = a.tpose; src[4] = a.ttwist; src[5] = a.mu; src[6] = a.normals; }
  else { src[0] = a.q; src[1] = a.quat; src[2] = a.wrench; src[3] = a.mu; src[4] = a.normals; src[5] = nullptr; src[6] = nullptr; }
  if (TMA) {
    if (lane == 0) {
      uint32_t bytes = 0;
#pragma unroll
      for (int s = 0; s < SG::kNumSeg; s++)
        if (src[s] != nullptr) bytes += SG::rows(s) * SG::kRow;
      mbar_expect_tx(bar, bytes);
      const uint32_t dst = smem_u32(stage);
#pragma unroll
      for (int s = 0; s < SG::kNumSeg; s++)
        if (src[s] != nullptr) tma_load_2d(dst + SG::offset(s), &maps.seg[s], (int)(tile * 8ull), 0, bar);
    }
  } else {
    const unsigned long long B = a.B;
    const int col = lane & 7;
    unsigned long long bx = tile * 8ull + col;
    if (bx >= B) bx = B - 1;
#pragma unroll
    for (int s = 0; s < SG::kNumSeg; s++) {
      if (src[s] == nullptr) continue;
      const uint32_t dst = smem_u32(stage) + SG::offset(s);
#pragma unroll
      for (int r0 = 0; r0 < SG::rows(s); r0 += 4) {
        const int r = r0 + (lane >> 3);
        if (r < SG::rows(s)) cp_async_elem(dst + (r * 8 + col) * (int)sizeof(real), src[s] + (size_t)r * B + bx);
      }
    }
    cp_async_commit();
  }
}

// The raw inputs of this lane's state from the staging buffer (the layout RawIn of qlb_solve_quad.cuh expects).
template <typename real, int MODE>
__device__ __forceinline__ void stage_read(const SolveArgsT<real>& a, const unsigned char* stage, const real mu_default,
                                           const int leg, const int quad, RawIn<real, MODE>& in) {
  using SG = Staging<real, MODE>;
  auto at = [&](const int seg, const int row) { return reinterpret_cast<const real*>(stage + SG::offset(seg))[row * 8 + quad]; };
#pragma unroll
  for (int j = 0; j < 3; j++) in.qj[j] = at(0, 3 * leg + j);
  if (MODE == 1) {
#pragma unroll
    for (int r = 0; r < 7; r++) { in.pose[r] = at(1, r); in.tp[r] = at(3, r); }
#pragma unroll
    for (int r = 0; r < 6; r++) { in.tw[r] = at(2, r); in.tt[r] = at(4, r); }
  } else {
#pragma unroll
    for (int r = 0; r < 4; r++) in.quat[r] = at(1, r);
#pragma unroll
    for (int r = 0; r < 6; r++) in.b[r] = at(2, r);
  }
  in.mu = (a.mu != nullptr) ? at(SG::kSegMu, leg) : mu_default;
  in.nw[0] = real(0.0); in.nw[1] = real(0.0); in.nw[2] = real(1.0);
  if (a.normals != nullptr) {
#pragma unroll
    for (int c = 0; c < 3; c++) in.nw[c] = at(SG::kSegNormals, 3 * leg + c);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Registers of one lane while its quad runs rounds on a hard state.
template <typename creal>
struct RoundState {
  creal At[3][6];   // the leg's block of the wrench map in contact coordinates (zero for a swing leg)
  creal mu;
  creal y[3];       // primal iterate (y_n, y_1, y_2)
  creal u[3];       // multipliers of the rows held by the slots n, 1, 2 (0 where the slot is free)
  int a0, sg1, sg2; // working set: y_n pinned at F_min; y_1 = sg1 mu y_n; y_2 = sg2 mu y_n
  float gscale;
  unsigned mask;
  bool alive;
  int rounds;
  unsigned idx;
};

// One round of the dual block active-set method for the states held by the eight quads (whole warp).
// b: the quad's desired wrench in shared memory (element r at b[r * bstride]).
// done: the iterate is optimal (KKT verified).  fail: factorisation failed or the round limit is reached.
template <typename creal>
__device__ __forceinline__ void dbas_round(RoundState<creal>& q, const creal* b, const int bstride, const CoreConst<creal>& cc,
                                           const int leg, const bool active, bool& done, bool& fail) {
  const bool alive = q.alive && active;
  const creal mu = q.mu;
  const creal (&At)[3][6] = q.At;
  const int a0 = q.a0, sg1 = q.sg1, sg2 = q.sg2;
  // ---- reduced columns of the equality-constrained QP for the working set
  creal v[3][6], al[3], r6[6];
  {
    const creal q1 = sg1 * mu, q2 = sg2 * mu;
    const bool fn = alive && a0 == 0, f1 = alive && sg1 == 0, f2 = alive && sg2 == 0;
    const creal wn = cc.W * fma(mu * mu, (creal)(sg1 * sg1 + sg2 * sg2), creal(1.0));
    al[0] = fn ? fast_rcp(wn) : creal(0.0);
    al[1] = f1 ? cc.winv : creal(0.0);
    al[2] = f2 ? cc.winv : creal(0.0);
    const creal pin = (alive && a0 != 0) ? cc.fmin : creal(0.0);
#pragma unroll
    for (int r = 0; r < 6; r++) {
      const creal cn = fma(q2, At[2][r], fma(q1, At[1][r], At[0][r]));
      v[0][r] = fn ? cn : creal(0.0);
      v[1][r] = f1 ? At[1][r] : creal(0.0);
      v[2][r] = f2 ? At[2][r] : creal(0.0);
      r6[r] = ((leg == 0 && active) ? b[r * bstride] : creal(0.0)) - pin * cn;
    }
  }
  // ---- the 6x6 system, summed over the quad
  creal N[21], rdg[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const creal w0 = al[0] * v[0][i], w1 = al[1] * v[1][i], w2 = al[2] * v[2][i];
#pragma unroll
    for (int j = 0; j < 6; j++) {
      if (j <= i) {
        creal acc = (i == j && leg == 0) ? cc.sinv[i] : creal(0.0);
        acc = fma(w0, v[0][j], acc);
        acc = fma(w1, v[1][j], acc);
        acc = fma(w2, v[2][j], acc);
        N[QLB_TRI(i, j)] = quad_sum(acc);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 6; r++) r6[r] = quad_sum(r6[r]);
  const bool pd = chol6_thread(N, rdg);
  if (Tol<creal>::refine) {
    creal rhs6[6];
#pragma unroll
    for (int r = 0; r < 6; r++) rhs6[r] = r6[r];
    solve6_thread(N, rdg, r6);
    refine6(N, rdg, v, al, cc.sinv, rhs6, r6);
  } else {
    solve6_thread(N, rdg, r6);   // r6 <- t = S (b - A y+)
  }
  // ---- y+ and the multipliers u+ of the working set
  creal zt[3], att[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    creal d0 = creal(0.0), d1 = creal(0.0);
#pragma unroll
    for (int r = 0; r < 6; r++) { d0 = fma(v[c][r], r6[r], d0); d1 = fma(At[c][r], r6[r], d1); }
    zt[c] = al[c] * d0;
    att[c] = d1;
  }
  const creal yn = (a0 != 0) ? cc.fmin : zt[0];
  creal yp[3];
  yp[0] = alive ? yn : creal(0.0);
  yp[1] = alive ? ((sg1 != 0) ? sg1 * mu * yn : zt[1]) : creal(0.0);
  yp[2] = alive ? ((sg2 != 0) ? sg2 * mu * yn : zt[2]) : creal(0.0);
  // gradient w y - A_k' t; stationarity g = D~' u gives the multipliers of the active rows
  const creal g0 = fma(cc.W, yp[0], -att[0]), g1 = fma(cc.W, yp[1], -att[1]), g2 = fma(cc.W, yp[2], -att[2]);
  creal up[3];
  up[1] = (alive && sg1 != 0) ? -sg1 * g1 : creal(0.0);
  up[2] = (alive && sg2 != 0) ? -sg2 * g2 : creal(0.0);
  up[0] = (alive && a0 != 0) ? g0 - mu * (up[1] + up[2]) : creal(0.0);
  // ---- ratio test: the step towards (y+, u+) that keeps every multiplier non-negative
  const creal tol_u = Tol<creal>::mult() * (creal)q.gscale;
  const bool act[3] = {alive && a0 != 0, alive && sg1 != 0, alive && sg2 != 0};
  creal tr[3], tl = creal(2.0);
#pragma unroll
  for (int s = 0; s < 3; s++) {
    tr[s] = creal(2.0);
    if (act[s] && up[s] < -tol_u) tr[s] = (q.u[s] > creal(0.0)) ? q.u[s] * fast_rcp(q.u[s] - up[s]) : creal(0.0);
    tl = fmin(tl, tr[s]);
  }
  creal tq = tl;
  tq = fmin(tq, __shfl_xor_sync(kFull, tq, 1));
  tq = fmin(tq, __shfl_xor_sync(kFull, tq, 2));
  const bool blocked = tq < creal(1.5);   // some multiplier turns negative on the way (tr <= 1 by construction)
  const creal step = blocked ? fmin(tq, creal(1.0)) : creal(1.0);
  // ---- move
  creal yn3[3], un3[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    yn3[c] = blocked ? fma(step, yp[c] - q.y[c], q.y[c]) : yp[c];
    un3[c] = blocked ? fmax(fma(step, up[c] - q.u[c], q.u[c]), creal(0.0)) : fmax(up[c], creal(0.0));
  }
  int na0 = a0, nsg1 = sg1, nsg2 = sg2;
  if (blocked) {
    // drop the blocking rows (all rows that share the smallest ratio: the fresh rows with a negative multiplier at t = 0)
    if (act[0] && tr[0] <= tq) { na0 = 0; un3[0] = creal(0.0); }
    if (act[1] && tr[1] <= tq) { nsg1 = 0; un3[1] = creal(0.0); }
    if (act[2] && tr[2] <= tq) { nsg2 = 0; un3[2] = creal(0.0); }
  }
  // ---- at the minimiser of the face: which rows does it violate?
  creal e[5];
  leg_rows(yn3[0], yn3[1], yn3[2], mu, e);
  e[0] -= cc.fmin;
  const float scale = fmaxf(1.f, quad_max(fmaxf(fabsf((float)yn3[0]), fmaxf(fabsf((float)yn3[1]), fabsf((float)yn3[2])))));
  const creal tol_s = Tol<creal>::feas() * (creal)scale;
  bool viol = false;
  if (!blocked && alive) {
    if (na0 == 0 && e[0] < -tol_s) { na0 = 1; viol = true; }
    if (nsg1 == 0) {
      const bool v1 = e[1] < -tol_s, v2 = e[2] < -tol_s;
      if (v1 || v2) { nsg1 = (v1 && (!v2 || e[1] <= e[2])) ? -1 : 1; viol = true; }
    }
    if (nsg2 == 0) {
      const bool v3 = e[3] < -tol_s, v4 = e[4] < -tol_s;
      if (v3 || v4) { nsg2 = (v3 && (!v4 || e[3] <= e[4])) ? -1 : 1; viol = true; }
    }
  }
  const bool quad_viol = quad_or(viol ? 1u : 0u) != 0u;
  done = false; fail = false;
  if (active) {
    q.rounds++;
    if (!pd) {
      fail = true;
    } else {
      q.y[0] = yn3[0]; q.y[1] = yn3[1]; q.y[2] = yn3[2];
      q.u[0] = un3[0]; q.u[1] = un3[1]; q.u[2] = un3[2];
      if (!blocked && !quad_viol) {
        done = true;            // the working set (q.a0, q.sg1, q.sg2) is the active set
      } else {
        q.a0 = na0; q.sg1 = nsg1; q.sg2 = nsg2;
        if (q.rounds >= QLB_DBAS_MAX_ROUNDS) fail = true;
      }
    }
  }
}

template <typename real, typename creal, int CAP>
__device__ __forceinline__ void stash_load(const WarpStash<real, creal, CAP>& ws, const int slot, const int leg, RoundState<creal>& q) {
  constexpr int Q = 4 * CAP;
  const int e = 4 * slot + leg;
  const uint32_t w1 = ws.sh[CAP + slot];
  q.idx = ws.sh[slot];
  q.gscale = __uint_as_float(ws.sh[2 * CAP + slot]);
  q.mask = w1 & 0xFu;
  q.alive = (w1 >> leg) & 1u;
  q.rounds = (int)(w1 >> 24);
  const unsigned pat = (w1 >> (4 + 5 * leg)) & 31u;
  q.a0 = (int)(pat & 1u);
  q.sg1 = ((pat >> 1) & 3u) == 1u ? -1 : (((pat >> 1) & 3u) == 2u ? 1 : 0);
  q.sg2 = ((pat >> 3) & 3u) == 1u ? -1 : (((pat >> 3) & 3u) == 2u ? 1 : 0);
  creal foot[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
#pragma unroll
    for (int k = 0; k < 3; k++) q.At[c][k] = ws.sl[(3 * c + k) * Q + e];
    foot[c] = ws.sl[(9 + c) * Q + e];
  }
  q.mu = ws.sl[12 * Q + e];
#pragma unroll
  for (int c = 0; c < 3; c++) { q.y[c] = ws.sl[(13 + c) * Q + e]; q.u[c] = ws.sl[(16 + c) * Q + e]; }
  // torque rows of the wrench map: r x e_c (zero for a swing leg: its force rows are stored as zero)
#pragma unroll
  for (int c = 0; c < 3; c++) {
    q.At[c][3] = foot[1] * q.At[c][2] - foot[2] * q.At[c][1];
    q.At[c][4] = foot[2] * q.At[c][0] - foot[0] * q.At[c][2];
    q.At[c][5] = foot[0] * q.At[c][1] - foot[1] * q.At[c][0];
  }
}

__device__ __forceinline__ unsigned pattern_bits(const int a0, const int sg1, const int sg2) {
  return (a0 != 0 ? 1u : 0u) | (sg1 == -1 ? 2u : (sg1 == 1 ? 4u : 0u)) | (sg2 == -1 ? 8u : (sg2 == 1 ? 16u : 0u));
}

// Write the iterate of an unfinished state back to its slot (the fixed part of the entry is still there).
template <typename real, typename creal, int CAP>
__device__ __forceinline__ void stash_save(const WarpStash<real, creal, CAP>& ws, const int slot, const int leg, const RoundState<creal>& q,
                                           const bool doit) {
  constexpr int Q = 4 * CAP;
  const unsigned pat = quad_or(pattern_bits(q.a0, q.sg1, q.sg2) << (5 * leg));   // whole warp
  if (doit) {
    const int e = 4 * slot + leg;
#pragma unroll
    for (int c = 0; c < 3; c++) { ws.sl[(13 + c) * Q + e] = q.y[c]; ws.sl[(16 + c) * Q + e] = q.u[c]; }
    if (leg == 0) ws.sh[CAP + slot] = q.mask | (pat << 4) | ((unsigned)q.rounds << 24);
  }
}

// The round phase of one warp (see the header).  occ: bit s = slot s holds a pending state.
template <typename real, typename creal, int CAP>
__device__ __forceinline__ void round_phase(const SolveArgsT<real>& a, const CoreConst<creal>& cc, const WarpStash<real, creal, CAP>& ws,
                                            unsigned& occ, const bool final, const int run_min, const int lane, const int leg,
                                            const int quad) {
  constexpr int Q = 4 * CAP;
  unsigned unassigned = occ;
  RoundState<creal> q;
#pragma unroll
  for (int c = 0; c < 3; c++) {
#pragma unroll
    for (int r = 0; r < 6; r++) q.At[c][r] = creal(0.0);
    q.y[c] = creal(0.0); q.u[c] = creal(0.0);
  }
  q.mu = creal(0.0); q.a0 = 0; q.sg1 = 0; q.sg2 = 0; q.gscale = 1.f; q.mask = 0u; q.alive = false; q.rounds = 0; q.idx = 0u;
  int slot = nth_set_bit(unassigned, quad);
  bool active = slot >= 0;
  {
    const int ntake = min(8, __popc(unassigned));
#pragma unroll 1
    for (int i = 0; i < ntake; i++) unassigned &= unassigned - 1u;   // the eight lowest pending slots are taken
  }
  if (active) stash_load(ws, slot, leg, q);
#pragma unroll 1
  for (;;) {
    bool done, fail;
    const int bslot = active ? slot : 0;
    dbas_round<creal>(q, ws.sb + bslot, CAP, cc, leg, active, done, fail);
    const bool leave = active && (done || fail);
    // ---- finished states: forces, torques, net wrench, flags (whole warp: quad shuffles inside)
    if (__any_sync(kFull, leave)) {
      LegSetup<creal> L;
#pragma unroll
      for (int c = 0; c < 3; c++) {
#pragma unroll
        for (int r = 0; r < 6; r++) L.At[c][r] = q.At[c][r];
      }
      L.alive = q.alive; L.mask = q.mask;
      const int oslot = active ? slot : 0;
      quad_output<real, creal>(a, L, q.y, q.a0, q.sg1, q.sg2, 0, q.rounds, (unsigned long long)q.idx, active && done, leg,
                               ws.sj + 4 * oslot + leg, Q);
      // not verified within the round limit (or a factorisation failed): the interior-point kernel takes the state
      const unsigned fm = __ballot_sync(kFull, active && fail && leg == 0);
      if (fm != 0u) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(a.list2_count, __popc(fm));
        base = __shfl_sync(kFull, base, 0);
        if (active && fail && leg == 0) a.list2[base + __popc(fm & ((1u << lane) - 1u))] = q.idx;
      }
      // release the slots, hand the next pending states to the quads that became free
      const unsigned freed = __reduce_or_sync(kFull, leave ? (1u << slot) : 0u);
      occ &= ~freed;
      const unsigned wm = __ballot_sync(kFull, leave && leg == 0);
      const int rank = __popc(wm & ((1u << (lane & ~3)) - 1u));
      __syncwarp();
      if (leave) {
        slot = nth_set_bit(unassigned, rank);
        active = slot >= 0;
        if (active) stash_load(ws, slot, leg, q);
      }
      {
        const int ntake = min(__popc(wm), __popc(unassigned));
#pragma unroll 1
        for (int i = 0; i < ntake; i++) unassigned &= unassigned - 1u;
      }
    }
    const int nact = __popc(__ballot_sync(kFull, active && leg == 0));
    if (nact == 0) break;
    if (!final && nact + __popc(unassigned) < run_min) {
      // too few states left to keep the warp busy: park the unfinished ones and fetch more tiles
      stash_save(ws, active ? slot : 0, leg, q, active);
      break;
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------
template <typename real, typename creal, int MODE, bool TMA>
__global__ void __launch_bounds__(kQuadThreads, QLB_FUSED_MIN_CTAS)
qlb_fused_kernel(const SolveArgsT<real> a, const __grid_constant__ FusedMaps maps) {
  using FL = FusedLayout<real, creal, MODE>;
  constexpr int CAP = FL::kCap;
  constexpr int Q = 4 * CAP;
  constexpr int kRunMin = CAP - 7;            // a tile needs eight free slots: the round phase runs down to CAP - 8 pending
  extern __shared__ __align__(128) unsigned char smem[];
  DeviceParamsT<real>& prm = *reinterpret_cast<DeviceParamsT<real>*>(smem);
  CoreConst<creal>& cc = *reinterpret_cast<CoreConst<creal>*>(smem + ((sizeof(DeviceParamsT<real>) + 15) & ~15));
  static_assert(((sizeof(DeviceParamsT<real>) + 15) & ~15) + sizeof(CoreConst<creal>) + 64 <= FL::kFixed, "fixed part");
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + FL::kFixed - 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int leg = lane & 3, quad = lane >> 2;
  unsigned char* wbase = smem + FL::kFixed + warp * FL::kWarpBytes;
  unsigned char* stage = wbase;
  const WarpStash<real, creal, CAP> ws(wbase + FL::kStage);
  const uint32_t bar = smem_u32(&bars[warp]);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.params);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&prm);
    for (int i = threadIdx.x; i < (int)(sizeof(DeviceParamsT<real>) / 4); i += blockDim.x) dst[i] = src[i];
    cc.load(a.params64);
    load_model_to_smem(a.model);
    if (TMA && lane == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
  __syncthreads();
  const unsigned long long B = a.B;
  const unsigned long long ntiles = (B + 7) / 8;
  const creal winv = cc.winv, cfmin = cc.fmin;
  unsigned occ = 0u;           // pending slots of the stash (warp-uniform)
  uint32_t parity = 0;

  auto claim = [&]() {
    unsigned long long b = 0;
    if (lane == 0) b = atomicAdd(a.counter, 1ull);
    return __shfl_sync(kFull, b, 0);
  };
  auto mask_of = [&](const unsigned long long tile) -> unsigned {
    const unsigned long long s = tile * 8ull + quad;
    return (tile < ntiles && s < B) ? (unsigned)a.mask[s] & 0xFu : 0u;
  };
  // two tiles are claimed ahead: the loads of `cur` are in flight, `nxt` is known, the claim after it is being fetched
  unsigned long long cur = claim();
  if (cur < ntiles) stage_issue<real, MODE, TMA>(a, maps, cur, stage, bar, lane);
  unsigned mask_cur = mask_of(cur);
  unsigned long long nxt = claim();
#pragma unroll 1
  for (;;) {
    const bool have = cur < ntiles;   // warp-uniform
    if (have) {
    const unsigned mask_nxt = mask_of(nxt);
    // ---- the staged rows of this tile
    if (TMA) { mbar_wait(bar, parity); parity ^= 1u; }
    else { cp_async_wait_all(); __syncwarp(); }
    RawIn<real, MODE> in;
    stage_read<real, MODE>(a, stage, prm.mu_default, leg, quad, in);
    const unsigned long long s0 = cur * 8ull + quad;
    const bool valid = s0 < B;
    const unsigned long long bq = valid ? s0 : (B - 1);
    in.mask = valid ? mask_cur : 0u;
    __syncwarp();     // every lane has read the buffer: the next tile may land in it
    if (nxt < ntiles) stage_issue<real, MODE, TMA>(a, maps, nxt, stage, bar, lane);
    const unsigned long long nn = claim();
    // ---- kinematics and QP data; the Jacobian goes straight into the slot this quad would keep
    const int slot = nth_set_bit(~occ & ((1u << CAP) - 1u), quad);
    LegSetup<creal> L;
    {
      LegSetup<real> L0;
      quad_setup<real, MODE>(a, prm, in, bq, valid, true, leg, L0, ws.sj + 4 * slot + leg, Q);
      widen_setup(L0, L);
    }
    int status;
    creal y[3], t[6];
    bool hard;
    unsigned pat;
    quad_first_solve<real, creal>(L, cc.sinv, winv, cfmin, leg, y, t, status, hard, pat);
    hard = hard && valid;
    creal net[6];
#pragma unroll
    for (int r = 0; r < 6; r++) net[r] = fma(-cc.sinv[r], t[r], L.b[r]);   // A x = b - S^-1 t
    // every state is written, the hard ones provisionally (full sectors; the round phase overwrites them while the
    // lines are still in L2)
    quad_output<real, creal>(a, L, y, 0, 0, 0, status, 0, bq, valid, leg, ws.sj + 4 * slot + leg, Q, net);
    // ---- park the hard states
    if (hard) {
      const int e = 4 * slot + leg;
#pragma unroll
      for (int c = 0; c < 3; c++) {
#pragma unroll
        for (int k = 0; k < 3; k++) ws.sl[(3 * c + k) * Q + e] = L.At[c][k];
        ws.sl[(9 + c) * Q + e] = L.foot[c];
        ws.sl[(13 + c) * Q + e] = y[c];
        ws.sl[(16 + c) * Q + e] = creal(0.0);
      }
      ws.sl[12 * Q + e] = L.mu;
      // lane `leg` stores components leg and leg + 4 (no dynamic register indexing)
      ws.sb[leg * CAP + slot] = (leg == 0) ? L.b[0] : (leg == 1 ? L.b[1] : (leg == 2 ? L.b[2] : L.b[3]));
      if (leg < 2) ws.sb[(4 + leg) * CAP + slot] = (leg == 0) ? L.b[4] : L.b[5];
      if (leg == 0) {
        ws.sh[slot] = (unsigned)bq;
        ws.sh[CAP + slot] = L.mask | (pat << 4);
        ws.sh[2 * CAP + slot] = __float_as_uint(L.gscale);
      }
    }
    occ |= __reduce_or_sync(kFull, hard ? (1u << slot) : 0u);
    __syncwarp();
    cur = nxt; nxt = nn; mask_cur = mask_nxt;
    }
    // one call site (the code of the round phase exists once): after a tile when the stash is full enough, and
    // once more when the tiles are exhausted, until the stash is empty
    const int pending = __popc(occ);
    if (have ? (pending >= kRunMin) : (pending > 0)) round_phase<real, creal, CAP>(a, cc, ws, occ, !have, kRunMin, lane, leg, quad);
    if (!have) break;
  }
}

}  // namespace qlb
