// rollout_dual.cuh — the warp-specialised rollout kernel with TWO rollout groups per leg warp (evaluation only).
//
// rollout_ws.cuh runs one group of 32 rollouts per CTA: 4 leg warps + 1 base warp, and per integrator sub-step the leg warps
// idle while the base warp solves the 6 x 6 system (a ~350-instruction dependent chain), the base warp while the legs run their
// phases.  4 CTAs per SM (the register file's limit at 96 registers) leave 22 % of the issue slots empty, and round 2 measured
// that neither re-phasing the resident CTAs nor any occupancy the register file allows changes that (profiles/README.md r2).
//
// Here a CTA holds two groups A and B = 64 rollouts: warps 0..3 = leg FL / FR / RL / RR of BOTH groups, warps 4 / 5 = the base
// role of group A / B.  A leg warp alternates "turns":
//        turn(X, i) = phase 2 of sub-step i of group X  ->  phase 1 of sub-step i + 1 of group X
// A, B, A, B, ...  While a leg warp works on B, group A's base warp solves and integrates, and vice versa: a leg warp never waits
// for a base solve that had a whole turn (~1 100 instructions) to finish.  Between its turns a group keeps, per rollout and leg,
// only the joint state (6 floats) and torques (3) in registers; the joint projections phase 2 needs (LegKeep: 35 floats) are
// parked in shared memory as float4 (9 LDS.128 + 9 STS.128 per turn, ~2 % of its instructions).  That is the interleaving the
// hardware scheduler provides between CTAs, at ~1.15x instead of 2x the registers per resident rollout: 3 CTAs x 64 rollouts per
// SM instead of 4 x 32, on 18 warps.
//
// Synchronisation: producer / consumer named barriers (bar.arrive for the producer, bar.sync for the consumer), three per group
// plus the two leg-pair barriers of rollout_ws.cuh:
//        A   legs arrive (pair sums in shared memory)        -> base syncs, solves, stores a0
//        B1  base arrives (a0)                               -> legs sync at the start of the group's next turn
//        B2  base arrives (R / v0 / pz of the new state)     -> legs sync before phase 1
// Every leg warp executes the same static schedule, so the pair barriers and the per-group barriers cannot deadlock: turn(X, i)
// only waits for work that was enabled by the end of turn(X, i - 1).
// The arithmetic is go2_ws.cuh's (costs equal rollout_ws_kernel's to 4e-8: ptxas contracts a few FMAs differently).
//
// MEASURED (r2, C = 4096, S = 1730, H = 5) and therefore NOT the default (SPI_B200_WS_DUAL=1 selects it): 51.9 ms at 3 CTAs per SM
// (96 registers, no spills, 12 leg + 6 base warps), 62.3 ms at 2 — against 41.4 ms for rollout_ws_kernel (16 leg + 4 base warps)
// and 46.5 ms for its 3-CTA build (12 + 3).  Leg warps that never wait for their base role are not faster: a leg warp issues
// ~0.2 instructions per cycle whatever it waits for in between (dependent FMA chains of a 3-joint recursion in 96 registers), so
// throughput follows the number of resident leg warps, 4 per sub-partition x 0.2 = the 78 % issue utilisation ncu reports, and
// the barrier stall is where that latency shows up, not its cause.  The lever is instruction-level parallelism per warp (or
// registers per SM), not the role schedule.
#pragma once
#include "rollout_ws.cuh"

namespace ws {

constexpr int kDualWarps = 6;
constexpr int kDualThreads = 32 * kDualWarps;
constexpr int kParkVec = 9;                        // LegKeep: 14 + 14 + 7 = 35 floats -> 9 float4

struct DualSmem {
  WsSmem g[2];
  float4 park[2][4][kParkVec][32];
};
constexpr int kDualSmemBytes = (int)sizeof(DualSmem);

__device__ __forceinline__ void dual_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void dual_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// barrier ids of group g (0 is left to __syncthreads)
__device__ __forceinline__ int bar_A(int g) { return 1 + 5 * g; }
__device__ __forceinline__ int bar_B1(int g) { return 2 + 5 * g; }
__device__ __forceinline__ int bar_B2(int g) { return 3 + 5 * g; }
__device__ __forceinline__ int bar_pair(int g, int leg) { return 4 + 5 * g + (leg >> 1); }

__device__ __forceinline__ void park_keep(float4 (*park)[32], int lane, const LegKeep& K) {
  park[0][lane] = make_float4(K.k1.cs, K.k1.sn, K.k1.cab, K.k1.cac);
  park[1][lane] = make_float4(K.k1.clb, K.k1.clc, K.k1.Ua[0], K.k1.Ua[1]);
  park[2][lane] = make_float4(K.k1.Ua[2], K.k1.Ul[0], K.k1.Ul[1], K.k1.Ul[2]);
  park[3][lane] = make_float4(K.k1.dinv, K.k1.u, K.k2.cs, K.k2.sn);
  park[4][lane] = make_float4(K.k2.cab, K.k2.cac, K.k2.clb, K.k2.clc);
  park[5][lane] = make_float4(K.k2.Ua[0], K.k2.Ua[1], K.k2.Ua[2], K.k2.Ul[0]);
  park[6][lane] = make_float4(K.k2.Ul[1], K.k2.Ul[2], K.k2.dinv, K.k2.u);
  park[7][lane] = make_float4(K.k3.cs, K.k3.sn, K.k3.cab, K.k3.cac);
  park[8][lane] = make_float4(K.k3.clb, K.k3.clc, K.k3.u, 0.f);
}
__device__ __forceinline__ void unpark_keep(const float4 (*park)[32], int lane, LegKeep& K) {
  float4 t;
  t = park[0][lane]; K.k1.cs = t.x; K.k1.sn = t.y; K.k1.cab = t.z; K.k1.cac = t.w;
  t = park[1][lane]; K.k1.clb = t.x; K.k1.clc = t.y; K.k1.Ua[0] = t.z; K.k1.Ua[1] = t.w;
  t = park[2][lane]; K.k1.Ua[2] = t.x; K.k1.Ul[0] = t.y; K.k1.Ul[1] = t.z; K.k1.Ul[2] = t.w;
  t = park[3][lane]; K.k1.dinv = t.x; K.k1.u = t.y; K.k2.cs = t.z; K.k2.sn = t.w;
  t = park[4][lane]; K.k2.cab = t.x; K.k2.cac = t.y; K.k2.clb = t.z; K.k2.clc = t.w;
  t = park[5][lane]; K.k2.Ua[0] = t.x; K.k2.Ua[1] = t.y; K.k2.Ua[2] = t.z; K.k2.Ul[0] = t.w;
  t = park[6][lane]; K.k2.Ul[1] = t.x; K.k2.Ul[2] = t.y; K.k2.dinv = t.z; K.k2.u = t.w;
  t = park[7][lane]; K.k3.cs = t.x; K.k3.sn = t.y; K.k3.cab = t.z; K.k3.cac = t.w;
  t = park[8][lane]; K.k3.clb = t.x; K.k3.clc = t.y; K.k3.u = t.z;
}

// what a leg warp keeps in registers for a group between its turns
struct DualCtx {
  LegState s;
  float tau[3];
  int c, seg;
};

struct DualGroup { int c, cta_in_cand, seg; bool active, live; };
__device__ __forceinline__ DualGroup dual_group(const WsArgs& A, int g, int lane) {
  DualGroup G;
  const long long n_groups = (long long)A.C_grid * A.n_cta_per_cand;
  long long grp = (long long)blockIdx.x * 2 + g;
  G.live = grp < n_groups;
  if (!G.live) grp = n_groups - 1;          // an odd group count: the last CTA's second group replays the last group, writes off
  const int cg = (int)(grp / A.n_cta_per_cand);
  G.cta_in_cand = (int)(grp - (long long)cg * A.n_cta_per_cand);
  const int seg_raw = G.cta_in_cand * kWsRollouts + lane;
  G.active = G.live && seg_raw < A.S;
  G.seg = seg_raw < A.S ? seg_raw : A.S - 1;
  G.c = cg;
  return G;
}

// PD torques of the physics step that starts at sub-step i (legged_robot_base.py:201-209: recomputed every physics step from
// the fresh joint state); the action, gains and motor parameters are re-read (L2-resident) instead of kept per group
__device__ __forceinline__ void dual_torques(const WsArgs& A, const LegK& L, int LEG, const DualCtx& X, int i, const int* motor_col,
                                             float* tau) {
  const SimK& S = A.M.sim;
  const int k = i / (S.nsub * A.decimation);
  const float* act_row = A.seg_actions + ((size_t)X.seg * A.H + k) * 12 + 3 * LEG;
  float act[3], kp[3], kd[3], motor[3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    act[j] = fminf(fmaxf(__ldg(act_row + j), -S.action_clip), S.action_clip);
    kp[j] = A.seg_gains ? __ldg(A.seg_gains + (size_t)X.seg * 24 + 3 * LEG + j) : A.M.kp[3 * LEG + j];
    kd[j] = A.seg_gains ? __ldg(A.seg_gains + (size_t)X.seg * 24 + 12 + 3 * LEG + j) : A.M.kd[3 * LEG + j];
    motor[j] = motor_col[j] >= 0 ? __ldg(A.params + (size_t)X.c * A.P + motor_col[j]) : 20.0f;
  }
  if (A.motor_model == SPI_MOTOR_SCALAR) { motor[1] = motor[0]; motor[2] = motor[0]; }
  leg_torques(S, L, act, X.s.q, X.s.qd, kp, kd, motor, A.motor_model, A.flags, tau);
}

// phase 1 of one sub-step of group g: state from the base -> outward / contact / inward pass -> pair sum -> signal the base
__device__ __forceinline__ void dual_phase1(const WsArgs& A, DualSmem& dsm, int g, int lane, int LEG, const LegK& L, DualCtx& X) {
  WsSmem& sm = dsm.g[g];
  dual_sync(bar_B2(g), kWsThreads);                 // the group's base role has published R / v0 / pz of this sub-step
  float bc[kBaseOut];
  ws_load_state(sm, lane, bc);
  float out[4 * kLegVec];
  out[4 * kLegVec - 1] = 0.f;
  LegKeep K;
  leg_phase1(A.M.sim, L, bc, X.s, X.tau, K, out, nullptr);
  park_keep(dsm.park[g][LEG], lane, K);
  if ((LEG & 1) == 0) {
#pragma unroll
    for (int v = 0; v < kLegVec; v++)
      sm.part[LEG][v][lane] = make_float4(out[4 * v], out[4 * v + 1], out[4 * v + 2], out[4 * v + 3]);
    dual_arrive(bar_pair(g, LEG), 64);
  } else {
    dual_sync(bar_pair(g, LEG), 64);
#pragma unroll
    for (int v = 0; v < kLegVec; v++) {
      const float4 p = sm.part[LEG - 1][v][lane];
      sm.part[LEG][v][lane] = make_float4(p.x + out[4 * v], p.y + out[4 * v + 1], p.z + out[4 * v + 2], p.w + out[4 * v + 3]);
    }
  }
  dual_arrive(bar_A(g), kWsThreads);                // the pair sums are in shared memory
}

__device__ __forceinline__ void dual_leg_role(const WsArgs& A, DualSmem& dsm, int lane, const int LEG) {
  const SimK& S = A.M.sim;
  const LegK& L = A.M.leg[LEG];
  // columns of this leg's motor parameters in the candidate rows (warp-uniform)
  int motor_col[3] = {-1, -1, -1};
  if (A.params)
    for (int p = 0; p < A.ids.n; p++) {
      const int id = A.ids.id[p];
      if (id == SPI_PARAM_MOTOR_HIP) motor_col[0] = p;
      else if (id == SPI_PARAM_MOTOR_THIGH) motor_col[1] = p;
      else if (id == SPI_PARAM_MOTOR_CALF) motor_col[2] = p;
    }
  const float h = S.dt / (float)S.nsub;
  const int n_sub = A.H * A.decimation * S.nsub;
  DualCtx cur, oth;
  // prologue: both groups' joint states, first torques, phase 1 of sub-step 0
#pragma unroll
  for (int g = 0; g < 2; g++) {
    DualCtx& X = g == 0 ? cur : oth;
    const DualGroup G = dual_group(A, g, lane);
    X.c = G.c; X.seg = G.seg;
    const float* row = A.seg_init + (size_t)G.seg * SPI_STATE_DIM;
#pragma unroll
    for (int j = 0; j < 3; j++) { X.s.q[j] = __ldg(row + 13 + 3 * LEG + j); X.s.qd[j] = __ldg(row + 25 + 3 * LEG + j); }
    dual_torques(A, L, LEG, X, 0, motor_col, X.tau);
    dual_phase1(A, dsm, g, lane, LEG, L, X);
  }
  // turns: A, B, A, B, ...   (turn t: group t & 1, sub-step t >> 1)
  int g = 0;
  for (int t = 0; t < 2 * n_sub; t++) {
    const int i = t >> 1;
    WsSmem& sm = dsm.g[g];
    dual_sync(bar_B1(g), kWsThreads);               // a0 of sub-step i (the base had the other group's whole turn for it)
    float bc[kBaseOut];
    ws_load_a0(sm, lane, bc + kBcA0);
    LegKeep K;
    unpark_keep(dsm.park[g][LEG], lane, K);
    leg_phase2(L, bc, K, cur.s, h);
    if (i + 1 < n_sub) {
      if ((i + 1) % S.nsub == 0) dual_torques(A, L, LEG, cur, i + 1, motor_col, cur.tau);
      dual_phase1(A, dsm, g, lane, LEG, L, cur);
    } else {
      // scripts/eval.py:292 — this leg's share of the squared joint-position error
      const float* tgt = A.seg_target + (size_t)cur.seg * SPI_TARGET_DIM;
      float ej = 0.f, acc = 0.f;
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const float e = cur.s.q[j] - __ldg(tgt + 7 + 3 * LEG + j);
        ej += e * e;
        acc += cur.s.q[j] * 0.f + cur.s.qd[j] * 0.f;
      }
      sm.ej[LEG][lane] = ej;
      sm.finite[LEG][lane] = finite_acc(acc) ? 1 : 0;
      dual_arrive(bar_A(g), kWsThreads);            // [C]
    }
    // hand the register context over to the other group
    DualCtx tmp = cur; cur = oth; oth = tmp;
    g ^= 1;
  }
}

__device__ __forceinline__ void dual_base_role(const WsArgs& A, DualSmem& dsm, int lane, int g) {
  const SimK& S = A.M.sim;
  WsSmem& sm = dsm.g[g];
  const DualGroup G = dual_group(A, g, lane);
  const int c = G.c, seg = G.seg;
  BaseInertia B;
  {
    float motor_unused[3];
    apply_candidate(A.M, A.params ? A.params + (size_t)c * A.P : nullptr, A.ids, A.flags, B, motor_unused);
  }
  BaseState s;
  {
    const float* row = A.seg_init + (size_t)seg * SPI_STATE_DIM;
#pragma unroll
    for (int i = 0; i < 3; i++) { s.p[i] = __ldg(row + i); s.v[i] = __ldg(row + 7 + i); s.w[i] = __ldg(row + 10 + i); }
#pragma unroll
    for (int i = 0; i < 4; i++) s.quat[i] = __ldg(row + 3 + i);
  }
  float bc[kBaseOut];
#pragma unroll
  for (int i = 0; i < 6; i++) bc[i] = 0.f;
  base_publish(s, bc);
  ws_store_state(sm, lane, bc);
  float pb[6];
  base_bias(B, bc, pb);
  dual_arrive(bar_B2(g), kWsThreads);               // state of sub-step 0
  const float h = S.dt / (float)S.nsub;
  const int n_sub = A.H * A.decimation * S.nsub;
  for (int i = 0; i < n_sub; i++) {
    dual_sync(bar_A(g), kWsThreads);                // the legs' pair sums of sub-step i
    float legsum[4 * kLegVec];
#pragma unroll
    for (int v = 0; v < kLegVec; v++) {
      const float4 p01 = sm.part[1][v][lane], p23 = sm.part[3][v][lane];
      legsum[4 * v] = p01.x + p23.x;
      legsum[4 * v + 1] = p01.y + p23.y;
      legsum[4 * v + 2] = p01.z + p23.z;
      legsum[4 * v + 3] = p01.w + p23.w;
    }
    float a0[6];
    base_solve(B, legsum, pb, a0);
    ws_store_a0(sm, lane, a0);
    dual_arrive(bar_B1(g), kWsThreads);
    // (same reason as in rollout_ws.cuh: keep the integration behind the signal)
    {
      const float4 t0 = ws_lds_volatile(&sm.bc[kBcStateVec][lane]), t1 = ws_lds_volatile(&sm.bc[kBcStateVec + 1][lane]);
      a0[0] = t0.x; a0[1] = t0.y; a0[2] = t0.z; a0[3] = t0.w; a0[4] = t1.x; a0[5] = t1.y;
    }
    base_advance(S, a0, s, h, bc);
    if (i + 1 < n_sub) {
      ws_store_state(sm, lane, bc);
      dual_arrive(bar_B2(g), kWsThreads);
      {
        const float4 t2 = ws_lds_volatile(&sm.bc[2][lane]), t3 = ws_lds_volatile(&sm.bc[3][lane]);
        bc[kBcV0] = t2.y; bc[kBcV0 + 1] = t2.z; bc[kBcV0 + 2] = t2.w; bc[kBcV0 + 3] = t3.x; bc[kBcV0 + 4] = t3.y; bc[kBcV0 + 5] = t3.z;
      }
      base_bias(B, bc, pb);
    }
  }
  // scripts/eval.py:287-292 — L2 errors of the final state
  const float* tgt = A.seg_target + (size_t)seg * SPI_TARGET_DIM;
  float ep = 0.f, eq = 0.f, acc = 0.f;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float e = s.p[i] - __ldg(tgt + i);
    ep += e * e;
    acc += s.p[i] * 0.f + s.v[i] * 0.f + s.w[i] * 0.f;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float e = s.quat[i] - __ldg(tgt + 3 + i);
    eq += e * e;
    acc += s.quat[i] * 0.f;
  }
  dual_sync(bar_A(g), kWsThreads);                  // [C] the legs' joint errors
  const float ej = (sm.ej[0][lane] + sm.ej[1][lane]) + (sm.ej[2][lane] + sm.ej[3][lane]);
  const bool ok = finite_acc(acc) && (sm.finite[0][lane] & sm.finite[1][lane] & sm.finite[2][lane] & sm.finite[3][lane]);
  float err[3] = {sqrtf(ep), sqrtf(eq), sqrtf(ej)};
  if (!ok && G.active) atomicOr(A.bad + c, 1);
  if (A.per_seg && G.active) {
    float* o = A.per_seg + ((size_t)c * A.S + seg) * 3;
    o[0] = err[0]; o[1] = err[1]; o[2] = err[2];
  }
  const bool counts = G.active && (A.seg_mask ? (A.seg_mask[seg] != 0) : true);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float v = counts ? err[i] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    err[i] = v;
  }
  if (lane == 0 && G.live) {
    float* o = A.partial + ((size_t)c * A.n_cta_per_cand + G.cta_in_cand) * 3;
    o[0] = err[0]; o[1] = err[1]; o[2] = err[2];
  }
}

template <int MINB>
__global__ void __launch_bounds__(kDualThreads, MINB) rollout_dual_kernel(const __grid_constant__ WsArgs A) {
  extern __shared__ __align__(16) unsigned char dual_smem_raw[];
  DualSmem& dsm = *reinterpret_cast<DualSmem*>(dual_smem_raw);
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp < 4) dual_leg_role(A, dsm, lane, warp);
  else dual_base_role(A, dsm, lane, warp - 4);
}

}  // namespace ws
