// spi_b200.cu — kernels + C-ABI of libspi_b200.so (see include/spi_b200.h for the contract and the
// reference call sites each entry point replaces).  sm_100a only; no torch types; no CPU fallback.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <map>
#include <mutex>
#include <vector>

#include "aba_leg.cuh"
#include "active_step.cuh"
#include "fim_tc.cuh"
#include "mlp_tc.cuh"
#include "rollout_ws.cuh"

using namespace spi;

// ================================================================================================
// kernels
// ================================================================================================
namespace {

constexpr int kThreads = 128;           // 4 warps = 32 rollouts per CTA
constexpr int kRolloutsPerCta = kThreads / 4;
constexpr int kRolloutsPerWarp = 8;

struct EvalArgs {
  const DeviceModel* model;
  const float* params; int C, P; ParamIds ids;
  const float* seg_init; const float* seg_actions; const float* seg_target; const float* seg_gains;
  const unsigned char* seg_mask;
  int S, H, decimation, motor_model; unsigned flags;
  int n_cta_per_cand, n_warp_per_cand;
  int paired;
  float* partial;      // [C][n_warp_per_cand][3]
  float* per_seg;      // [C][S][3] or null
  int* bad;            // [C]
  float* out_states;   // [C][S][H][37] (RECORD)
  const unsigned char* zero_mask;   // paired mode: [S] 1 = action replaced by 0, or null
};

SPI_DEV void load_leg_const(const DeviceModel& M, int leg, LegConst& L) {
  const LegConst& G = M.leg[leg];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    L.m[j] = G.m[j];
    L.qdef[j] = G.qdef[j];
    L.tlim[j] = G.tlim[j];
    L.foot[j] = G.foot[j];
#pragma unroll
    for (int k = 0; k < 3; k++) { L.h[j][k] = G.h[j][k]; L.r[j][k] = G.r[j][k]; }
#pragma unroll
    for (int k = 0; k < 6; k++) L.Io[j][k] = G.Io[j][k];
  }
}

// plain loads: sim_step_kernel and the in-place env step write the rows they read (no ld.global.nc on such memory)
SPI_DEV void load_lane_state(const float* row, int leg, LaneState& s) {
#pragma unroll
  for (int i = 0; i < 3; i++) { s.p[i] = *(row + i); s.v[i] = *(row + 7 + i); s.w[i] = *(row + 10 + i); }
#pragma unroll
  for (int i = 0; i < 4; i++) s.quat[i] = *(row + 3 + i);
#pragma unroll
  for (int j = 0; j < 3; j++) { s.q[j] = *(row + 13 + 3 * leg + j); s.qd[j] = *(row + 25 + 3 * leg + j); }
}

SPI_DEV void store_lane_state(float* row, int leg, const LaneState& s) {
  if (leg == 0) {
#pragma unroll
    for (int i = 0; i < 3; i++) { row[i] = s.p[i]; row[7 + i] = s.v[i]; row[10 + i] = s.w[i]; }
#pragma unroll
    for (int i = 0; i < 4; i++) row[3 + i] = s.quat[i];
  }
#pragma unroll
  for (int j = 0; j < 3; j++) { row[13 + 3 * leg + j] = s.q[j]; row[25 + 3 * leg + j] = s.qd[j]; }
}

SPI_DEV bool lane_finite(const LaneState& s) {
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 3; i++) acc += s.p[i] * 0.f + s.v[i] * 0.f + s.w[i] * 0.f + s.q[i] * 0.f + s.qd[i] * 0.f;
#pragma unroll
  for (int i = 0; i < 4; i++) acc += s.quat[i] * 0.f;
  return acc == 0.f;  // NaN/Inf * 0 = NaN
}

// The fused hot path: reset -> H x (clip, decimation x (PD + motor model, nsub x ABA sub-step)) -> errors
// -> per-warp masked partial sums.  4 lanes per (candidate, segment) rollout.
template <bool RECORD, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) rollout_kernel(const EvalArgs A) {
  const DeviceModel& M = *A.model;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int leg = lane & 3;
  const int cg = blockIdx.x / A.n_cta_per_cand;
  const int cta_in_cand = blockIdx.x - cg * A.n_cta_per_cand;
  const int seg_raw = cta_in_cand * kRolloutsPerCta + (threadIdx.x >> 2);
  const bool active = seg_raw < A.S;
  const int seg = active ? seg_raw : A.S - 1;
  const int c = A.paired ? seg : cg;

  SimConst S = M.sim;
  LegConst L;
  load_leg_const(M, leg, L);
  BaseInertia B;
  float motor[3];
  {
    float motor_all[3];
    apply_candidate(M, A.params ? A.params + (size_t)c * A.P : nullptr, A.ids, A.flags, B, motor_all);
    // act2tau_scalar uses one gain for every joint; vec3 variants are per joint group = per joint of the leg
    motor[0] = motor_all[0]; motor[1] = motor_all[1]; motor[2] = motor_all[2];
    if (A.motor_model == SPI_MOTOR_SCALAR) { motor[1] = motor_all[0]; motor[2] = motor_all[0]; }
  }
  LaneState s;
  load_lane_state(A.seg_init + (size_t)seg * SPI_STATE_DIM, leg, s);
  float kp[3], kd[3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    kp[j] = A.seg_gains ? __ldg(A.seg_gains + (size_t)seg * 24 + 3 * leg + j) : M.kp[3 * leg + j];
    kd[j] = A.seg_gains ? __ldg(A.seg_gains + (size_t)seg * 24 + 12 + 3 * leg + j) : M.kd[3 * leg + j];
  }
  const float h = S.dt / (float)S.nsub;
  const float* act_row = A.seg_actions + (size_t)seg * A.H * 12 + 3 * leg;
  const bool zero_act = A.zero_mask && A.zero_mask[seg] != 0;
  for (int k = 0; k < A.H; k++) {
    float act[3];
#pragma unroll
    for (int j = 0; j < 3; j++)
      act[j] = zero_act ? 0.f : fminf(fmaxf(__ldg(act_row + 12 * k + j), -S.action_clip), S.action_clip);
    for (int d = 0; d < A.decimation; d++) {
      float tau[3];
      lane_torques(S, L, act, s.q, s.qd, kp, kd, motor, A.motor_model, A.flags, tau);
      for (int n = 0; n < S.nsub; n++) substep(S, L, B, s, tau, h, nullptr);
    }
    if (RECORD) {
      if (active) store_lane_state(A.out_states + (((size_t)(A.paired ? 0 : c) * A.S + seg) * A.H + k) * SPI_STATE_DIM, leg, s);
    }
  }
  if (RECORD) return;

  // scripts/eval.py:287-292 — L2 errors of the final state
  const float* tgt = A.seg_target + (size_t)seg * SPI_TARGET_DIM;
  float ep = 0.f, eq = 0.f, ej = 0.f;
#pragma unroll
  for (int i = 0; i < 3; i++) { const float e = s.p[i] - __ldg(tgt + i); ep += e * e; }
#pragma unroll
  for (int i = 0; i < 4; i++) { const float e = s.quat[i] - __ldg(tgt + 3 + i); eq += e * e; }
#pragma unroll
  for (int j = 0; j < 3; j++) { const float e = s.q[j] - __ldg(tgt + 7 + 3 * leg + j); ej += e * e; }
  ej = group_sum(ej);
  float err[3] = {sqrtf(ep), sqrtf(eq), sqrtf(ej)};
  bool ok = lane_finite(s);
  if (!ok && active) atomicOr(A.bad + c, 1);
  if (A.per_seg && active && leg == 0) {
    float* o = A.per_seg + ((size_t)c * A.S + seg) * 3;
    o[0] = err[0]; o[1] = err[1]; o[2] = err[2];
  }
  const bool counts = active && (A.seg_mask ? (A.seg_mask[seg] != 0) : true);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float v = counts ? err[i] : 0.f;
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    err[i] = v;
  }
  if (lane == 0) {
    float* o = A.partial + ((size_t)c * A.n_warp_per_cand + cta_in_cand * (kThreads / 32) + warp) * 3;
    o[0] = err[0]; o[1] = err[1]; o[2] = err[2];
  }
}

// fixed-order sum of the per-warp partials -> mean costs (scripts/eval.py:304-309)
__global__ void reduce_cost_kernel(const float* partial, const int* bad, const unsigned char* seg_mask, int C, int S,
                                   int n_warp_per_cand, float cost_denominator, float* out_cost, int* out_status) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * 3) return;
  const int c = idx / 3, k = idx - 3 * c;
  float denom = cost_denominator;
  if (!(denom > 0.f)) {
    int n = 0;
    if (seg_mask) { for (int s = 0; s < S; s++) n += seg_mask[s] ? 1 : 0; } else n = S;
    denom = (float)n;
  }
  float sum = 0.f;
  const float* p = partial + (size_t)c * n_warp_per_cand * 3 + k;
  for (int w = 0; w < n_warp_per_cand; w++) sum += p[3 * w];
  const int b = bad[c];
  out_cost[idx] = b ? INFINITY : sum / denom;
  if (out_status && k == 0) out_status[c] = b;
}

// BaseSimulator.simulate_at_each_physics_step for N envs (4 lanes per env), n_steps physics steps
struct StepArgs {
  const DeviceModel* model;
  const float* params; int P; ParamIds ids; unsigned flags;
  float* state; const float* torques; int N, n_steps;
  float* foot_force;
  const float* ext_wrench;   // [N,13,6] or null
};
__global__ void __launch_bounds__(kThreads) sim_step_kernel(const StepArgs A) {
  const DeviceModel& M = *A.model;
  const int leg = threadIdx.x & 3;
  const int env_raw = blockIdx.x * kRolloutsPerCta + (threadIdx.x >> 2);
  const bool active = env_raw < A.N;
  const int env = active ? env_raw : A.N - 1;
  SimConst S = M.sim;
  LegConst L;
  load_leg_const(M, leg, L);
  BaseInertia B;
  float motor[3];
  apply_candidate(M, A.params ? A.params + (size_t)env * A.P : nullptr, A.ids, A.flags, B, motor);
  LaneState s;
  load_lane_state(A.state + (size_t)env * SPI_STATE_DIM, leg, s);
  float tau[3], ff[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 3; j++) tau[j] = A.torques[(size_t)env * 12 + 3 * leg + j];
  const float h = S.dt / (float)S.nsub;
  float ext[18], ext_base[6];
  if (A.ext_wrench) {
    const float* w = A.ext_wrench + (size_t)env * 13 * 6;
#pragma unroll
    for (int i = 0; i < 6; i++) ext_base[i] = w[i];
#pragma unroll
    for (int i = 0; i < 18; i++) ext[i] = w[6 + 18 * leg + i];
  }
  for (int k = 0; k < A.n_steps; k++)
    for (int n = 0; n < S.nsub; n++)
      substep(S, L, B, s, tau, h, ff, A.ext_wrench ? ext : nullptr, A.ext_wrench ? ext_base : nullptr);
  if (active) {
    store_lane_state(A.state + (size_t)env * SPI_STATE_DIM, leg, s);
    if (A.foot_force) {
      float* o = A.foot_force + ((size_t)env * 4 + leg) * 3;
      o[0] = ff[0]; o[1] = ff[1]; o[2] = ff[2];
    }
  }
}

// ---- rigid-body state tensor by forward kinematics: one thread per (env, chain), chain 0..3 = legs, 4 = base + heads ------
struct FkModel {
  float r[4][3][3];      // joint origins hip / thigh / calf in the parent frame
  float foot[4][3];      // foot link origin in the calf frame
  float head[2][3];      // Head_upper / Head_lower link origins in the base frame
};
SPI_DEV void quat_mul(const float* a, const float* b, float* o) {   // xyzw
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  o[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}
SPI_DEV void quat_rot(const float* q, const float* v, float* o) {   // R(q) v
  const float tx = 2.f * (q[1] * v[2] - q[2] * v[1]), ty = 2.f * (q[2] * v[0] - q[0] * v[2]), tz = 2.f * (q[0] * v[1] - q[1] * v[0]);
  o[0] = v[0] + q[3] * tx + (q[1] * tz - q[2] * ty);
  o[1] = v[1] + q[3] * ty + (q[2] * tx - q[0] * tz);
  o[2] = v[2] + q[3] * tz + (q[0] * ty - q[1] * tx);
}
SPI_DEV void fk_store(float* o, const float* p, const float* q, const float* v, const float* w) {
  o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = q[0]; o[4] = q[1]; o[5] = q[2]; o[6] = q[3];
  o[7] = v[0]; o[8] = v[1]; o[9] = v[2]; o[10] = w[0]; o[11] = w[1]; o[12] = w[2];
}
// child link across a joint: origin offset r (parent frame), rotation by angle about `axis` (0 = x, 1 = y, -1 = fixed)
SPI_DEV void fk_child(const float* pp, const float* pq, const float* pv, const float* pw, const float* r, int axis,
                      float ang, float rate, float* cp, float* cq, float* cv, float* cw) {
  float rw[3];
  quat_rot(pq, r, rw);
  cp[0] = pp[0] + rw[0]; cp[1] = pp[1] + rw[1]; cp[2] = pp[2] + rw[2];
  cv[0] = pv[0] + (pw[1] * rw[2] - pw[2] * rw[1]);
  cv[1] = pv[1] + (pw[2] * rw[0] - pw[0] * rw[2]);
  cv[2] = pv[2] + (pw[0] * rw[1] - pw[1] * rw[0]);
  if (axis < 0) {
    cq[0] = pq[0]; cq[1] = pq[1]; cq[2] = pq[2]; cq[3] = pq[3];
    cw[0] = pw[0]; cw[1] = pw[1]; cw[2] = pw[2];
    return;
  }
  float sh, ch;
  sincosf(0.5f * ang, &sh, &ch);
  const float jq[4] = {axis == 0 ? sh : 0.f, axis == 1 ? sh : 0.f, 0.f, ch};
  quat_mul(pq, jq, cq);
  const float ax[3] = {axis == 0 ? rate : 0.f, axis == 1 ? rate : 0.f, 0.f};
  float aw[3];
  quat_rot(cq, ax, aw);
  cw[0] = pw[0] + aw[0]; cw[1] = pw[1] + aw[1]; cw[2] = pw[2] + aw[2];
}
__global__ void body_states_kernel(const FkModel F, const float* state, int N, float* out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * 5) return;
  const int e = idx / 5, chain = idx - 5 * e;
  const float* s = state + (size_t)e * SPI_STATE_DIM;
  float* o = out + (size_t)e * 19 * 13;
  const float bp[3] = {s[0], s[1], s[2]}, bq[4] = {s[3], s[4], s[5], s[6]}, bv[3] = {s[7], s[8], s[9]},
              bw[3] = {s[10], s[11], s[12]};
  if (chain == 4) {
    fk_store(o, bp, bq, bv, bw);
    for (int h = 0; h < 2; h++) {
      float p[3], q[4], v[3], w[3];
      fk_child(bp, bq, bv, bw, F.head[h], -1, 0.f, 0.f, p, q, v, w);
      fk_store(o + (9 + h) * 13, p, q, v, w);
    }
    return;
  }
  const int first = (chain < 2) ? 1 + 4 * chain : 11 + 4 * (chain - 2);   // FL 1, FR 5, RL 11, RR 15
  float pp[3] = {bp[0], bp[1], bp[2]}, pq[4] = {bq[0], bq[1], bq[2], bq[3]}, pv[3] = {bv[0], bv[1], bv[2]},
        pw[3] = {bw[0], bw[1], bw[2]};
  for (int j = 0; j < 4; j++) {
    float p[3], q[4], v[3], w[3];
    if (j < 3) fk_child(pp, pq, pv, pw, F.r[chain][j], j == 0 ? 0 : 1, s[13 + 3 * chain + j], s[25 + 3 * chain + j], p, q, v, w);
    else fk_child(pp, pq, pv, pw, F.foot[chain], -1, 0.f, 0.f, p, q, v, w);
    fk_store(o + (first + j) * 13, p, q, v, w);
    for (int k = 0; k < 3; k++) { pp[k] = p[k]; pv[k] = v[k]; pw[k] = w[k]; }
    for (int k = 0; k < 4; k++) pq[k] = q[k];
  }
}

// torque law on its own, one thread per (row, joint)
__global__ void torque_kernel(const DeviceModel* Mp, const float* actions, const float* q, const float* qd,
                              const float* gains, const float* motor_params, int N, int motor_model, unsigned flags,
                              float* out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * 12) return;
  const DeviceModel& M = *Mp;
  const int e = idx / 12, j = idx - 12 * e, leg = j / 3, jj = j - 3 * leg;
  const float clipv = M.sim.action_clip;
  float as = fminf(fmaxf(actions[idx], -clipv), clipv) * M.sim.action_scale;
  if (jj == 0 && (flags & SPI_FLAG_HIP_HALF)) as *= 0.5f;
  const float kp = gains ? gains[e * 24 + j] : M.kp[j];
  const float kd = gains ? gains[e * 24 + 12 + j] : M.kd[j];
  const float lim = M.leg[leg].tlim[jj];
  float t = kp * (as + M.leg[leg].qdef[jj] - q[idx]) - kd * qd[idx];
  const float g = motor_params ? motor_params[e * 3 + jj] : 20.0f;
  const float g0 = motor_params ? motor_params[e * 3] : 20.0f;
  if (motor_model == SPI_MOTOR_VEC3_TANH && (flags & SPI_FLAG_TANH_BEFORE_CLIP)) {
    t = g * tanhf((1.0f / g) * t);
    t = fminf(fmaxf(t, -lim), lim);
  } else {
    t = fminf(fmaxf(t, -lim), lim);
    if (motor_model == SPI_MOTOR_SCALAR) t *= g0;
    else if (motor_model == SPI_MOTOR_VEC3) t *= g;
    else if (motor_model == SPI_MOTOR_VEC3_TANH) t = g * tanhf((1.0f / g) * t);
  }
  out[idx] = t;
}

// Fisher-information reward (active_sysid_openloop.py:402-426), one warp per main env:
//   J[p][d] = (main[d] - aux_p[d]) / delta,  trace = sum J^2,  JJt[p][q] = sum_d J[p][d] J[q][d]
__global__ void fim_reward_kernel(const float* states, int Mn, int P, float delta, int accumulate, float* out_JtJ,
                                  float* out_trace) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Mn) return;
  const float* base = states + (size_t)warp * (P + 1) * 25;
  const float inv = 1.0f / delta;
  const float main_d = (lane < 25) ? base[lane] : 0.f;
  float tr = 0.f;
  for (int p = 0; p < P; p++) {
    const float jp = (lane < 25) ? (main_d - base[(size_t)(p + 1) * 25 + lane]) * inv : 0.f;
    tr += jp * jp;
    if (out_JtJ) {
      for (int q = 0; q <= p; q++) {
        const float jq = (lane < 25) ? (main_d - base[(size_t)(q + 1) * 25 + lane]) * inv : 0.f;
        float v = jp * jq;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
          float* d1 = out_JtJ + ((size_t)warp * P + p) * P + q;
          float* d2 = out_JtJ + ((size_t)warp * P + q) * P + p;
          if (accumulate) { *d1 += v; if (p != q) *d2 += v; }
          else { *d1 = v; *d2 = v; }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
  if (lane == 0 && out_trace) { if (accumulate) out_trace[warp] += tr; else out_trace[warp] = tr; }
}

// total[c] = w . cost[c,:]
__global__ void weighted_cost_kernel(const float* cost3, int C, float w0, float w1, float w2, float* out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = w0 * cost3[3 * c] + w1 * cost3[3 * c + 1] + w2 * cost3[3 * c + 2];
}

// ---- CEM on device -----------------------------------------------------------------------------
// counter-based RNG (Philox-4x32-10), keyed by (seed), counter = (global candidate, param, iteration)
SPI_DEV void philox4x32(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned* out) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1;
    const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__global__ void cem_sample_kernel(const float* mean, const float* stdv, const float* lo, const float* hi, int C, int P,
                                  int c0, unsigned long long seed, int iteration, float* out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * P) return;
  const int c = idx / P, p = idx - c * P;
  unsigned r[4];
  philox4x32((unsigned)(c0 + c), (unsigned)p, (unsigned)iteration, 0x5B1u, (unsigned)seed, (unsigned)(seed >> 32), r);
  const float u1 = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float z = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
  float v = mean[p] + stdv[p] * z;
  if (lo) v = fmaxf(v, lo[p]);
  if (hi) v = fminf(v, hi[p]);
  out[idx] = v;
}
// Elite selection in O(C): the n_elite smallest candidates in the stable order (cost, index), NaN / inf last — the same set
// the O(C^2) rank kernel of round 1 produced (1.07e9 compares at 32768 candidates, replicated on every rank: it was the
// first thing to show in multi-GPU scaling).  One CTA: 4 passes of an 8-bit radix select over order-preserving uint keys
// find the n_elite-th smallest key K*; every key < K* is elite, ties at K* are admitted in index order until n_elite is
// reached (block-wide exclusive scan over contiguous index chunks).  Output in the old kernel's vocabulary so that the
// refit kernel is unchanged: rank[i] = 0 for the best candidate, 1 for the other elites, n_elite for the rest.
__device__ __forceinline__ unsigned cem_key(float c) {
  if (!(fabsf(c) <= 3.4028235e38f)) c = INFINITY;    // NaN and +-inf sort last (cem.cem_refit_numpy: non-finite last)
  const unsigned b = __float_as_uint(c);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
constexpr int kSelThreads = 1024;
__global__ void __launch_bounds__(kSelThreads) cem_select_kernel(const float* cost, int C, int n_elite, int* rank) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_remaining, s_best_key;
  __shared__ int s_best_idx;
  __shared__ unsigned scan[kSelThreads];
  const int t = threadIdx.x;
  if (t == 0) { s_prefix = 0u; s_remaining = (unsigned)n_elite; s_best_key = 0xFFFFFFFFu; s_best_idx = 0x7FFFFFFF; }
  __syncthreads();
  // ---- radix select of the n_elite-th smallest key (1-based position s_remaining among the keys matching s_prefix) ----
  for (int pass = 0; pass < 4; pass++) {
    const int shift = 24 - 8 * pass;
    const unsigned mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    if (t < 256) hist[t] = 0u;
    __syncthreads();
    const unsigned prefix = s_prefix;
    for (int i = t; i < C; i += kSelThreads) {
      const unsigned k = cem_key(cost[i]);
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    if (t == 0) {
      unsigned rem = s_remaining, d = 0;
      for (; d < 256; d++) { if (hist[d] >= rem) break; rem -= hist[d]; }
      s_prefix = prefix | (d << shift);
      s_remaining = rem;                               // position among the keys equal to the new prefix
    }
    __syncthreads();
  }
  const unsigned kstar = s_prefix;
  const unsigned ties_admitted = s_remaining;          // how many keys == K* are elite (>= 1)
  // ---- ties in index order: contiguous chunk per thread, exclusive scan of the per-chunk tie counts ----
  const int chunk = (C + kSelThreads - 1) / kSelThreads;
  const int i0 = min(t * chunk, C), i1 = min(i0 + chunk, C);
  unsigned n_tie = 0;
  unsigned best_key = 0xFFFFFFFFu; int best_idx = 0x7FFFFFFF;
  for (int i = i0; i < i1; i++) {
    const unsigned k = cem_key(cost[i]);
    n_tie += (k == kstar) ? 1u : 0u;
    if (k < best_key) { best_key = k; best_idx = i; }  // first index wins within the chunk
  }
  scan[t] = n_tie;
  __syncthreads();
  for (int o = 1; o < kSelThreads; o <<= 1) {          // Hillis-Steele inclusive scan
    const unsigned v = (t >= o) ? scan[t - o] : 0u;
    __syncthreads();
    scan[t] += v;
    __syncthreads();
  }
  unsigned tie_before = scan[t] - n_tie;
  // ---- the best candidate: minimum (key, index) ----
  atomicMin(&s_best_key, best_key);
  __syncthreads();
  if (best_key == s_best_key) atomicMin(&s_best_idx, best_idx);
  __syncthreads();
  const int best = s_best_idx;
  for (int i = i0; i < i1; i++) {
    const unsigned k = cem_key(cost[i]);
    bool elite = k < kstar;
    if (k == kstar) { elite = tie_before < ties_admitted; tie_before++; }
    rank[i] = (i == best) ? 0 : (elite ? 1 : n_elite);
  }
}
// one CTA per parameter: elite mean / std in a fixed summation order, smoothed update; CTA 0 also
// writes the best candidate (rank 0)
__global__ void cem_refit_kernel(const float* params, const float* cost, const int* rank, int C, int P, int n_elite,
                                 float alpha, const float* std_floor, float* mean, float* stdv, float* out_best) {
  __shared__ float sh[256];
  __shared__ float sh_mean;
  const int p = blockIdx.x, t = threadIdx.x;
  float acc = 0.f;
  for (int i = t; i < C; i += 256) acc += (rank[i] < n_elite) ? params[(size_t)i * P + p] : 0.f;   // rank: 0 best, 1 elite, n_elite otherwise
  sh[t] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (t < o) sh[t] += sh[t + o]; __syncthreads(); }
  if (t == 0) sh_mean = sh[0] / (float)n_elite;
  __syncthreads();
  const float mu = sh_mean;
  acc = 0.f;
  for (int i = t; i < C; i += 256) {
    const float d = params[(size_t)i * P + p] - mu;
    acc += (rank[i] < n_elite) ? d * d : 0.f;
  }
  __syncthreads();
  sh[t] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (t < o) sh[t] += sh[t + o]; __syncthreads(); }
  if (t == 0) {
    float sd = sqrtf(sh[0] / (float)n_elite);
    float m2 = (1.f - alpha) * mean[p] + alpha * mu;
    float s2 = (1.f - alpha) * stdv[p] + alpha * sd;
    if (std_floor) s2 = fmaxf(s2, std_floor[p]);
    mean[p] = m2; stdv[p] = s2;
  }
  if (out_best) {
    for (int i = t; i < C; i += 256)
      if (rank[i] == 0) { out_best[p] = params[(size_t)i * P + p]; if (p == 0) out_best[P] = cost[i]; }
  }
}

// FP32 FMA peak: 8 independent FFMA chains per thread, register resident
__global__ void __launch_bounds__(256) fp32_peak_kernel(int iters, float seed, float* sink) {
  float a0 = seed, a1 = seed + 1.f, a2 = seed + 2.f, a3 = seed + 3.f, a4 = seed + 4.f, a5 = seed + 5.f, a6 = seed + 6.f,
        a7 = seed + 7.f;
  const float m = 0.999f, b = 0.001f;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      a0 = fmaf(a0, m, b); a1 = fmaf(a1, m, b); a2 = fmaf(a2, m, b); a3 = fmaf(a3, m, b);
      a4 = fmaf(a4, m, b); a5 = fmaf(a5, m, b); a6 = fmaf(a6, m, b); a7 = fmaf(a7, m, b);
    }
  }
  const float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (r == 123456.789f) sink[0] = r;
}

}  // namespace

// ================================================================================================
// host side: model handle, error handling, C-ABI
// ================================================================================================
struct spi_b200_model {
  FkModel fk;
  DeviceModel host_model;
  DeviceModel* d_model = nullptr;
  ws::ModelK ws_model;      // constants of the warp-specialised fast path (passed as a kernel parameter)
  bool ws_ok = false;       // the blob has the Go2-family structure the fast path is compiled for
  int kernel = SPI_KERNEL_AUTO;
  int device = 0;
  int sm_count = 148;
  // workspaces (grown on demand)
  float* d_partial = nullptr; size_t partial_cap = 0;
  int* d_bad = nullptr; size_t bad_cap = 0;
  int* d_rank = nullptr; size_t rank_cap = 0;
  // partial sums of the step slices of spi_b200_fim_contract, one buffer PER STREAM: pipelined explorers contract on their own
  // streams through the same handle, and a shared buffer would race
  struct FimPart { float* ptr = nullptr; size_t cap = 0; };
  std::map<cudaStream_t, FimPart> fim_part;
  std::mutex fim_part_mu;
  // staging for the *_host entry point
  char* d_stage = nullptr; size_t d_stage_cap = 0;
  char* h_stage = nullptr; size_t h_stage_cap = 0;
  // roofline instrumentation: event pairs around every rollout-kernel launch (spi_b200_timing_*)
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool;   // recycled events
  std::vector<cudaEvent_t> ev_pending;  // start0, stop0, start1, stop1, ...
  double timing_ms = 0.0; long long timing_launches = 0;
};

namespace {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
#define CUDA_OK(expr)                                                                             \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fail(-100, std::string(#expr) + ": " + cudaGetErrorString(_e));                      \
  } while (0)

// cudaFuncSetAttribute is PER DEVICE: a process that drives several GPUs (one engine per device) must opt in on each of them,
// so the "already done" state is a bitmap over the current device's index, not a process-wide flag.
static bool attr_needed_on_current_device(std::atomic<unsigned long long>* done_mask, int* dev_out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { *dev_out = -1; return true; }
  *dev_out = dev;
  return ((done_mask->load() >> dev) & 1ull) == 0;
}
static void attr_mark_done(std::atomic<unsigned long long>* done_mask, int dev) {
  if (dev >= 0) done_mask->fetch_or(1ull << dev);
}

int check_launch(const char* what) {
  g_launches.fetch_add(1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(-101, std::string(what) + " launch failed: " + cudaGetErrorString(e));
  return 0;
}

template <class T> int ensure(T** ptr, size_t* cap, size_t need) {
  if (*cap >= need) return 0;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr; *cap = 0;
  CUDA_OK(cudaMalloc((void**)ptr, need * sizeof(T)));
  *cap = need;
  return 0;
}

int make_ids(int P, const int* param_ids, ParamIds* out) {
  if (P < 0 || P > 16) return fail(-3, "P must be in [0,16]");
  out->n = P;
  for (int i = 0; i < 16; i++) out->id[i] = -1;
  for (int i = 0; i < P; i++) {
    if (!param_ids) return fail(-3, "param_ids is NULL");
    if (param_ids[i] < 0 || param_ids[i] >= SPI_PARAM_COUNT) return fail(-3, "param id out of range");
    out->id[i] = param_ids[i];
  }
  return 0;
}

int blob_to_model(const float* b, int n, DeviceModel* M) {
  if (!b || n < SPI_BLOB_SIZE) return fail(-2, "model blob too short");
  if (b[SPI_BLOB_MAGIC] != SPI_BLOB_MAGIC_VALUE) return fail(-2, "model blob magic mismatch");
  std::memset(M, 0, sizeof(*M));
  M->sim.dt = b[SPI_BLOB_DT]; M->sim.gz = b[SPI_BLOB_GRAVITY_Z];
  M->sim.action_scale = b[SPI_BLOB_ACTION_SCALE]; M->sim.action_clip = b[SPI_BLOB_ACTION_CLIP];
  M->sim.kn = b[SPI_BLOB_CONTACT_KN]; M->sim.cn = b[SPI_BLOB_CONTACT_CN]; M->sim.mu = b[SPI_BLOB_CONTACT_MU];
  M->sim.dtan = b[SPI_BLOB_CONTACT_DT]; M->sim.radius = b[SPI_BLOB_FOOT_RADIUS];
  M->sim.veps2 = b[SPI_BLOB_CONTACT_VEPS] * b[SPI_BLOB_CONTACT_VEPS];
  M->sim.nsub = (int)b[SPI_BLOB_NSUB];
  if (M->sim.nsub < 1) M->sim.nsub = 1;
  if (!(M->sim.dt > 0.f)) return fail(-2, "model blob: dt must be positive");
  for (int k = 0; k < 10; k++) M->base_inertial[k] = b[SPI_BLOB_BASE_INERTIAL + k];
  if (!(M->base_inertial[0] > 0.f)) return fail(-2, "model blob: base mass must be positive");
  for (int l = 0; l < 2; l++)
    for (int k = 0; k < 10; k++) M->lumps[l][k] = b[SPI_BLOB_BASE_LUMPS + 10 * l + k];
  for (int leg = 0; leg < 4; leg++) {
    LegConst& L = M->leg[leg];
    for (int j = 0; j < 3; j++) {
      const float* p = b + SPI_BLOB_LEG_BODIES + SPI_LEG_BODY_STRIDE * (3 * leg + j);
      const int axis = (int)p[13];
      // the leg-per-lane kernel is specialised for the Go2 chain hip(x) - thigh(y) - calf(y)
      if (axis != (j == 0 ? 0 : 1)) return fail(-2, "model blob: joint axes must be hip=x, thigh=y, calf=y");
      const float m = p[0];
      const float c[3] = {p[1], p[2], p[3]};
      const float cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
      L.m[j] = m;
      for (int k = 0; k < 3; k++) { L.h[j][k] = m * c[k]; L.r[j][k] = p[10 + k]; }
      const float Ic[6] = {p[4], p[5], p[6], p[7], p[8], p[9]};
      for (int r = 0; r < 3; r++)
        for (int k = r; k < 3; k++) {
          const int s = (r == k) ? r : ((r + k == 1) ? 3 : ((r + k == 2) ? 4 : 5));
          L.Io[j][s] = Ic[s] + m * ((r == k ? cc : 0.f) - c[r] * c[k]);
        }
      L.qdef[j] = b[SPI_BLOB_Q_DEFAULT + 3 * leg + j];
      L.tlim[j] = b[SPI_BLOB_TORQUE_LIMIT + 3 * leg + j];
    }
    for (int k = 0; k < 3; k++) L.foot[k] = b[SPI_BLOB_FOOT_OFFSET + 3 * leg + k];
  }
  for (int j = 0; j < 12; j++) { M->kp[j] = b[SPI_BLOB_KP + j]; M->kd[j] = b[SPI_BLOB_KD + j]; }
  return 0;
}

int timing_events(spi_b200_model* m, cudaEvent_t* e0, cudaEvent_t* e1) {
  cudaEvent_t* out[2] = {e0, e1};
  for (auto* e : out) {
    if (!m->ev_pool.empty()) { *e = m->ev_pool.back(); m->ev_pool.pop_back(); }
    else CUDA_OK(cudaEventCreate(e));
  }
  return 0;
}

// drain the pending event pairs into the accumulator (synchronises on each stop event)
int timing_drain(spi_b200_model* m) {
  for (size_t i = 0; i + 1 < m->ev_pending.size(); i += 2) {
    float ms = 0.f;
    CUDA_OK(cudaEventSynchronize(m->ev_pending[i + 1]));
    CUDA_OK(cudaEventElapsedTime(&ms, m->ev_pending[i], m->ev_pending[i + 1]));
    m->timing_ms += ms; m->timing_launches += 1;
    m->ev_pool.push_back(m->ev_pending[i]); m->ev_pool.push_back(m->ev_pending[i + 1]);
  }
  m->ev_pending.clear();
  return 0;
}

// SPI_B200_KERNEL = "lane" forces the generic leg-per-lane kernel; default = warp-specialised fast path when the
// model qualifies.  SPI_B200_MINB = min CTAs per SM the kernels are compiled for (occupancy experiments).
int kernel_choice() {
  static const int v = [] { const char* e = getenv("SPI_B200_KERNEL"); return (e && std::string(e) == "lane") ? 1 : 0; }();
  return v;
}
// SPI_B200_PDL=0 disables programmatic dependent launch of the closed-loop step's kernels (pdl.cuh); default on
bool pdl_on() {
  static const int v = [] { const char* e = getenv("SPI_B200_PDL"); return e ? atoi(e) : 1; }();
  return v != 0;
}
int minb_choice() {
  static const int v = [] { const char* e = getenv("SPI_B200_MINB"); return e ? atoi(e) : 0; }();
  return v;
}

#if defined(SPI_WS_PROFILE)   // dev builds only: tools/ws_timeline.py
long long* g_ws_prof = nullptr; int g_ws_prof_blk0 = 0, g_ws_prof_nblk = 0;
#endif

int launch_rollout_ws(spi_b200_model* m, bool record, const float* params, int C, int P, const int* param_ids,
                      const float* seg_init, const float* seg_actions, const float* seg_target, const float* seg_gains,
                      const unsigned char* seg_mask, int S, int H, int decimation, int motor_model, unsigned flags,
                      float cost_denominator, float* out_cost, float* out_per_seg, int* out_status, float* out_states,
                      cudaStream_t st, int paired, const unsigned char* zero_mask) {
  ws::WsArgs A;
  std::memset(&A, 0, sizeof(A));
  ParamIds ids;
  if (int rc = make_ids(P, param_ids, &ids)) return rc;
  A.ids.n = ids.n;
  for (int i = 0; i < 16; i++) A.ids.id[i] = ids.id[i];
  A.M = m->ws_model;
  A.params = (P > 0) ? params : nullptr; A.C = C; A.P = P;
  A.seg_init = seg_init; A.seg_actions = seg_actions; A.seg_target = seg_target; A.seg_gains = seg_gains;
  A.seg_mask = seg_mask; A.S = S; A.H = H; A.decimation = decimation; A.motor_model = motor_model; A.flags = flags;
  A.n_cta_per_cand = (S + ws::kWsRollouts - 1) / ws::kWsRollouts;
  A.paired = paired; A.zero_mask = zero_mask; A.C_grid = C;
  // the throughput launch packs the left-over segments (S % 32) of several candidates into shared tail CTAs (WsArgs::tail_lanes)
  static const int dense_env = [] { const char* e = getenv("SPI_B200_WS_DENSE"); return e ? atoi(e) : 1; }();
  const int n_full = S / ws::kWsRollouts, left = S - n_full * ws::kWsRollouts;
  A.tail_lanes = 0;
  if (!record && !paired && dense_env && m->kernel != SPI_KERNEL_WS_PADDED && left > 0) {
    int lp = 1;
    while (lp < left) lp <<= 1;
    if (lp < ws::kWsRollouts) A.tail_lanes = lp;        // (left > 16: a tail CTA would hold one candidate — the padded launch)
  }
  { static const int rot = getenv("SPI_B200_WS_ROT") ? atoi(getenv("SPI_B200_WS_ROT")) : 0; A.rotate_roles = rot; }
  const int cand_per_tail = A.tail_lanes ? ws::kWsRollouts / A.tail_lanes : 1;
  const long long n_cta = A.tail_lanes ? (long long)C * n_full + (C + cand_per_tail - 1) / cand_per_tail : (long long)C * A.n_cta_per_cand;
  if (n_cta > 2147483647LL) return fail(-3, "C * ceil(S/32) exceeds the grid limit");
#if defined(SPI_WS_PROFILE)
  A.prof = g_ws_prof; A.prof_blk0 = g_ws_prof_blk0; A.prof_nblk = g_ws_prof_nblk;
#endif
  const int minb = minb_choice();
  if (record) {
    A.out_states = out_states;
    // 2 CTAs per SM (no register cap) is the faster kernel per CTA; once the grid does not fit in one wave of it (296 CTAs:
    // e.g. the 352 CTAs of a config-5 control step) the 4-CTAs-per-SM build keeps the step in a single wave
    static const int rec_minb = [] { const char* e = getenv("SPI_B200_RECORD_MINB"); return e ? atoi(e) : 0; }();
    const bool four = rec_minb ? (rec_minb == 4) : (n_cta > 2LL * m->sm_count);
    // (programmatic dependent launch: in the closed-loop step this kernel follows the actor's last layer on the same stream)
    if (four) CUDA_OK(pdl::launch(pdl_on(), ws::rollout_ws_kernel<true, 4>, dim3((unsigned)n_cta), dim3(ws::kWsThreads), 0, st, A));
    else CUDA_OK(pdl::launch(pdl_on(), ws::rollout_ws_kernel<true, 2>, dim3((unsigned)n_cta), dim3(ws::kWsThreads), 0, st, A));
    return check_launch("rollout_ws_kernel<record>");
  }
  if (int rc = ensure(&m->d_partial, &m->partial_cap, (size_t)C * A.n_cta_per_cand * 3)) return rc;
  if (int rc = ensure(&m->d_bad, &m->bad_cap, (size_t)C)) return rc;
  CUDA_OK(cudaMemsetAsync(m->d_bad, 0, (size_t)C * sizeof(int), st));
  A.partial = m->d_partial; A.bad = m->d_bad; A.per_seg = out_per_seg;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (m->timing) {
    if (int rc = timing_events(m, &e0, &e1)) return rc;
    CUDA_OK(cudaEventRecord(e0, st));
  }
  if (minb == 2 || (minb == 0 && n_cta <= 2LL * m->sm_count))
    // small grids (a mass_opt trial is 55 CTAs, the 20-point landscape 1100 -> no: only grids that fit 2 CTAs per SM): the
    // uncapped-register build has the shorter per-CTA latency, and latency is all a sub-wave launch has
    ws::rollout_ws_kernel<false, 2><<<(unsigned)n_cta, ws::kWsThreads, 0, st>>>(A);
  else if (minb == 3) ws::rollout_ws_kernel<false, 3><<<(unsigned)n_cta, ws::kWsThreads, 0, st>>>(A);
  else if (minb == 5) ws::rollout_ws_kernel<false, 5><<<(unsigned)n_cta, ws::kWsThreads, 0, st>>>(A);
  else {
    // the throughput kernel, one instance per motor model (rollout_ws.cuh: MOTOR = 2 * model + tanh-before-clip)
    const int motor = 2 * motor_model + ((motor_model == SPI_MOTOR_VEC3_TANH && (flags & SPI_FLAG_TANH_BEFORE_CLIP)) ? 1 : 0);
    const unsigned g = (unsigned)n_cta;
    switch (motor) {
      case 2 * SPI_MOTOR_NONE: ws::rollout_ws_kernel<false, 4, 2 * SPI_MOTOR_NONE><<<g, ws::kWsThreads, 0, st>>>(A); break;
      case 2 * SPI_MOTOR_SCALAR: ws::rollout_ws_kernel<false, 4, 2 * SPI_MOTOR_SCALAR><<<g, ws::kWsThreads, 0, st>>>(A); break;
      case 2 * SPI_MOTOR_VEC3: ws::rollout_ws_kernel<false, 4, 2 * SPI_MOTOR_VEC3><<<g, ws::kWsThreads, 0, st>>>(A); break;
      case 2 * SPI_MOTOR_VEC3_TANH: ws::rollout_ws_kernel<false, 4, 2 * SPI_MOTOR_VEC3_TANH><<<g, ws::kWsThreads, 0, st>>>(A); break;
      case 2 * SPI_MOTOR_VEC3_TANH + 1: ws::rollout_ws_kernel<false, 4, 2 * SPI_MOTOR_VEC3_TANH + 1><<<g, ws::kWsThreads, 0, st>>>(A); break;
      default: ws::rollout_ws_kernel<false, 4><<<g, ws::kWsThreads, 0, st>>>(A); break;
    }
  }
  if (int rc = check_launch("rollout_ws_kernel")) return rc;
  if (m->timing) {
    CUDA_OK(cudaEventRecord(e1, st));
    m->ev_pending.push_back(e0); m->ev_pending.push_back(e1);
  }
  const int n = C * 3;
  reduce_cost_kernel<<<(n + 127) / 128, 128, 0, st>>>(m->d_partial, m->d_bad, seg_mask, C, S, A.n_cta_per_cand,
                                                      cost_denominator, out_cost, out_status);
  return check_launch("reduce_cost_kernel");
}

int launch_rollout(spi_b200_model* m, bool record, const float* params, int C, int P, const int* param_ids,
                   const float* seg_init, const float* seg_actions, const float* seg_target, const float* seg_gains,
                   const unsigned char* seg_mask, int S, int H, int decimation, int motor_model, unsigned flags,
                   float cost_denominator, float* out_cost, float* out_per_seg, int* out_status, float* out_states,
                   cudaStream_t st, int paired = 0, const unsigned char* zero_mask = nullptr) {
  if (!m) return fail(-1, "model handle is NULL");
  if (C <= 0 || S <= 0 || H <= 0 || decimation <= 0) return fail(-3, "C, S, H, decimation must be positive");
  if (!seg_init || !seg_actions) return fail(-3, "seg_init / seg_actions is NULL");
  if (!record && (!seg_target || !out_cost)) return fail(-3, "seg_target / out_cost is NULL");
  if (record && !out_states) return fail(-3, "out_states is NULL");
  if (P > 0 && !params) return fail(-3, "params is NULL");
  if (motor_model < SPI_MOTOR_NONE || motor_model > SPI_MOTOR_VEC3_TANH) return fail(-3, "unknown motor_model");
  if (m->kernel >= SPI_KERNEL_WS && !m->ws_ok) return fail(-6, "the model blob does not have the Go2-family structure of the fast path");
  if (m->ws_ok && m->kernel != SPI_KERNEL_LANE && (m->kernel >= SPI_KERNEL_WS || kernel_choice() != 1))
    return launch_rollout_ws(m, record, params, C, P, param_ids, seg_init, seg_actions, seg_target, seg_gains, seg_mask,
                             S, H, decimation, motor_model, flags, cost_denominator, out_cost, out_per_seg, out_status,
                             out_states, st, paired, zero_mask);
  EvalArgs A;
  std::memset(&A, 0, sizeof(A));
  if (int rc = make_ids(P, param_ids, &A.ids)) return rc;
  A.model = m->d_model;
  A.params = (P > 0) ? params : nullptr; A.C = C; A.P = P;
  A.seg_init = seg_init; A.seg_actions = seg_actions; A.seg_target = seg_target; A.seg_gains = seg_gains;
  A.seg_mask = seg_mask; A.S = S; A.H = H; A.decimation = decimation; A.motor_model = motor_model; A.flags = flags;
  A.n_cta_per_cand = (S + kRolloutsPerCta - 1) / kRolloutsPerCta;
  A.n_warp_per_cand = A.n_cta_per_cand * (kThreads / 32);
  A.paired = paired; A.zero_mask = zero_mask;
  const long long n_cta = (long long)C * A.n_cta_per_cand;
  if (n_cta > 2147483647LL) return fail(-3, "C * ceil(S/32) exceeds the grid limit");
  if (!record) {
    if (int rc = ensure(&m->d_partial, &m->partial_cap, (size_t)C * A.n_warp_per_cand * 3)) return rc;
    if (int rc = ensure(&m->d_bad, &m->bad_cap, (size_t)C)) return rc;
    CUDA_OK(cudaMemsetAsync(m->d_bad, 0, (size_t)C * sizeof(int), st));
    A.partial = m->d_partial; A.bad = m->d_bad; A.per_seg = out_per_seg;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (m->timing) {
      if (int rc = timing_events(m, &e0, &e1)) return rc;
      CUDA_OK(cudaEventRecord(e0, st));
    }
    const int variant = minb_choice();
    if (variant == 3) rollout_kernel<false, 3><<<(unsigned)n_cta, kThreads, 0, st>>>(A);
    else if (variant == 4) rollout_kernel<false, 4><<<(unsigned)n_cta, kThreads, 0, st>>>(A);
    else rollout_kernel<false, 2><<<(unsigned)n_cta, kThreads, 0, st>>>(A);
    if (int rc = check_launch("rollout_kernel")) return rc;
    if (m->timing) {
      CUDA_OK(cudaEventRecord(e1, st));
      m->ev_pending.push_back(e0); m->ev_pending.push_back(e1);
    }
    const int n = C * 3;
    reduce_cost_kernel<<<(n + 127) / 128, 128, 0, st>>>(m->d_partial, m->d_bad, seg_mask, C, S, A.n_warp_per_cand,
                                                        cost_denominator, out_cost, out_status);
    return check_launch("reduce_cost_kernel");
  }
  A.out_states = out_states;
  rollout_kernel<true, 2><<<(unsigned)n_cta, kThreads, 0, st>>>(A);
  return check_launch("rollout_kernel<record>");
}

}  // namespace

extern "C" {

int spi_b200_version(void) { return SPI_B200_VERSION; }
const char* spi_b200_last_error(void) { return g_last_error.c_str(); }
long long spi_b200_launch_count(void) { return g_launches.load(); }

int spi_b200_model_create(const float* model_blob, int n_floats, spi_b200_model** out_model) {
  if (!out_model) return fail(-1, "out_model is NULL");
  *out_model = nullptr;
  spi_b200_model* m = new (std::nothrow) spi_b200_model();
  if (!m) return fail(-4, "out of host memory");
  if (int rc = blob_to_model(model_blob, n_floats, &m->host_model)) { delete m; return rc; }
  std::memset(&m->ws_model, 0, sizeof(m->ws_model));
  m->ws_ok = ws::model_from_blob(model_blob, &m->ws_model) == 0;
  for (int leg = 0; leg < 4; leg++) {
    for (int j = 0; j < 3; j++)
      for (int k = 0; k < 3; k++) m->fk.r[leg][j][k] = model_blob[SPI_BLOB_LEG_BODIES + SPI_LEG_BODY_STRIDE * (3 * leg + j) + 10 + k];
    for (int k = 0; k < 3; k++) m->fk.foot[leg][k] = model_blob[SPI_BLOB_FOOT_OFFSET + 3 * leg + k] - model_blob[SPI_BLOB_FOOT_SPHERE + k];
  }
  for (int h = 0; h < 2; h++)   // the head links' inertial frames coincide with their link frames in the URDF
    for (int k = 0; k < 3; k++) m->fk.head[h][k] = model_blob[SPI_BLOB_BASE_LUMPS + 10 * h + 1 + k];
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    delete m;
    return fail(-5, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0"));
  }
  e = cudaGetDevice(&m->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, m->device);
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_model, sizeof(DeviceModel));
  if (e == cudaSuccess) e = cudaMemcpy(m->d_model, &m->host_model, sizeof(DeviceModel), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (m->d_model) cudaFree(m->d_model);
    delete m;
    return fail(-100, std::string("model upload failed: ") + cudaGetErrorString(e));
  }
  *out_model = m;
  return 0;
}

int spi_b200_model_destroy(spi_b200_model* m) {
  if (!m) return 0;
  if (m->d_model) cudaFree(m->d_model);
  if (m->d_partial) cudaFree(m->d_partial);
  if (m->d_bad) cudaFree(m->d_bad);
  if (m->d_rank) cudaFree(m->d_rank);
  for (auto& kv : m->fim_part) if (kv.second.ptr) cudaFree(kv.second.ptr);
  if (m->d_stage) cudaFree(m->d_stage);
  if (m->h_stage) cudaFreeHost(m->h_stage);
  for (auto e : m->ev_pool) cudaEventDestroy(e);
  for (auto e : m->ev_pending) cudaEventDestroy(e);
  delete m;
  return 0;
}

int spi_b200_model_set_kernel(spi_b200_model* m, int kernel) {
  if (!m) return fail(-1, "model handle is NULL");
  if (kernel < SPI_KERNEL_AUTO || kernel > SPI_KERNEL_WS_PADDED) return fail(-3, "unknown kernel id");
  if (kernel >= SPI_KERNEL_WS && !m->ws_ok) return fail(-6, "the model blob does not have the Go2-family structure of the fast path");
  m->kernel = kernel;
  return 0;
}

int spi_b200_timing_enable(spi_b200_model* m, int enable) {
  if (!m) return fail(-1, "model handle is NULL");
  m->timing = enable != 0;
  return 0;
}

int spi_b200_timing_read(spi_b200_model* m, double* out_total_ms, long long* out_launches, int reset) {
  if (!m) return fail(-1, "model handle is NULL");
  if (int rc = timing_drain(m)) return rc;
  if (out_total_ms) *out_total_ms = m->timing_ms;
  if (out_launches) *out_launches = m->timing_launches;
  if (reset) { m->timing_ms = 0.0; m->timing_launches = 0; }
  return 0;
}

int spi_b200_eval_candidates(spi_b200_model* model, const float* params, int C, int P, const int* param_ids,
                             const float* seg_init, const float* seg_actions, const float* seg_target,
                             const float* seg_gains, const unsigned char* seg_mask, int S, int H, int decimation,
                             int motor_model, unsigned flags, float cost_denominator, float* out_cost,
                             float* out_per_seg, int* out_status, void* cuda_stream) {
  return launch_rollout(model, false, params, C, P, param_ids, seg_init, seg_actions, seg_target, seg_gains, seg_mask,
                        S, H, decimation, motor_model, flags, cost_denominator, out_cost, out_per_seg, out_status,
                        nullptr, (cudaStream_t)cuda_stream);
}

int spi_b200_rollout_states(spi_b200_model* model, const float* params, int C, int P, const int* param_ids,
                            const float* seg_init, const float* seg_actions, const float* seg_gains, int S, int H,
                            int decimation, int motor_model, unsigned flags, float* out_states, void* cuda_stream) {
  return launch_rollout(model, true, params, C, P, param_ids, seg_init, seg_actions, nullptr, seg_gains, nullptr, S, H,
                        decimation, motor_model, flags, 0.f, nullptr, nullptr, nullptr, out_states,
                        (cudaStream_t)cuda_stream);
}

int spi_b200_eval_candidates_host(spi_b200_model* m, const float* params, int C, int P, const int* param_ids,
                                  const float* seg_init, const float* seg_actions, const float* seg_target,
                                  const float* seg_gains, const unsigned char* seg_mask, int S, int H, int decimation,
                                  int motor_model, unsigned flags, float cost_denominator, float* out_cost,
                                  int* out_status, void* cuda_stream) {
  if (!m) return fail(-1, "model handle is NULL");
  if (C <= 0 || S <= 0 || H <= 0) return fail(-3, "C, S, H must be positive");
  if (!seg_init || !seg_actions || !seg_target || !out_cost) return fail(-3, "NULL host buffer");
  if (P > 0 && !params) return fail(-3, "params is NULL");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_params = al((size_t)C * P * 4), b_init = al((size_t)S * SPI_STATE_DIM * 4),
               b_act = al((size_t)S * H * 12 * 4), b_tgt = al((size_t)S * SPI_TARGET_DIM * 4),
               b_gain = seg_gains ? al((size_t)S * 24 * 4) : 0, b_mask = seg_mask ? al((size_t)S) : 0,
               b_cost = al((size_t)C * 3 * 4), b_stat = al((size_t)C * 4);
  const size_t in_bytes = b_params + b_init + b_act + b_tgt + b_gain + b_mask;
  const size_t total = in_bytes + b_cost + b_stat;
  if (m->d_stage_cap < total) {
    if (m->d_stage) cudaFree(m->d_stage);
    if (m->h_stage) cudaFreeHost(m->h_stage);
    m->d_stage = nullptr; m->h_stage = nullptr; m->d_stage_cap = m->h_stage_cap = 0;
    CUDA_OK(cudaMalloc((void**)&m->d_stage, total));
    CUDA_OK(cudaMallocHost((void**)&m->h_stage, total));
    m->d_stage_cap = m->h_stage_cap = total;
  }
  size_t o = 0;
  auto put = [&](const void* src, size_t bytes, size_t padded) -> size_t {
    size_t at = o;
    if (src && bytes) std::memcpy(m->h_stage + at, src, bytes);
    o += padded;
    return at;
  };
  const size_t o_params = put(params, (size_t)C * P * 4, b_params);
  const size_t o_init = put(seg_init, (size_t)S * SPI_STATE_DIM * 4, b_init);
  const size_t o_act = put(seg_actions, (size_t)S * H * 12 * 4, b_act);
  const size_t o_tgt = put(seg_target, (size_t)S * SPI_TARGET_DIM * 4, b_tgt);
  const size_t o_gain = put(seg_gains, (size_t)S * 24 * 4, b_gain);
  const size_t o_mask = put(seg_mask, (size_t)S, b_mask);
  const size_t o_cost = o, o_stat = o + b_cost;
  CUDA_OK(cudaMemcpyAsync(m->d_stage, m->h_stage, in_bytes, cudaMemcpyHostToDevice, st));
  char* d = m->d_stage;
  int rc = launch_rollout(m, false, (const float*)(d + o_params), C, P, param_ids, (const float*)(d + o_init),
                          (const float*)(d + o_act), (const float*)(d + o_tgt),
                          seg_gains ? (const float*)(d + o_gain) : nullptr,
                          seg_mask ? (const unsigned char*)(d + o_mask) : nullptr, S, H, decimation, motor_model, flags,
                          cost_denominator, (float*)(d + o_cost), nullptr, (int*)(d + o_stat), nullptr, st);
  if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(m->h_stage + o_cost, d + o_cost, b_cost + b_stat, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  std::memcpy(out_cost, m->h_stage + o_cost, (size_t)C * 3 * 4);
  if (out_status) std::memcpy(out_status, m->h_stage + o_stat, (size_t)C * 4);
  return 0;
}

int spi_b200_env_step(spi_b200_model* m, const float* params, int P, const int* param_ids, float* state,
                      const float* actions, const unsigned char* zero_action_mask, const float* gains, int N,
                      int decimation, int motor_model, unsigned flags, void* cuda_stream) {
  if (!m) return fail(-1, "model handle is NULL");
  if (N <= 0 || decimation <= 0) return fail(-3, "N and decimation must be positive");
  if (!state || !actions) return fail(-3, "state / actions is NULL");
  // one control step of N independent envs = the RECORD rollout kernel with H = 1 in paired mode (env e uses
  // parameter row e and state row e); every lane reads its row before it writes it, so in-place is safe
  return launch_rollout(m, true, params, 1, params ? P : 0, param_ids, state, actions, nullptr, gains, nullptr, N, 1,
                        decimation, motor_model, flags, 0.f, nullptr, nullptr, nullptr, state, (cudaStream_t)cuda_stream, 1,
                        zero_action_mask);
}

int spi_b200_sim_step(spi_b200_model* m, const float* params, int P, const int* param_ids, unsigned flags, float* state,
                      const float* torques, int N, int n_steps, float* out_foot_force, void* cuda_stream) {
  return spi_b200_sim_step_ext(m, params, P, param_ids, flags, state, torques, nullptr, N, n_steps, out_foot_force, cuda_stream);
}

int spi_b200_sim_step_ext(spi_b200_model* m, const float* params, int P, const int* param_ids, unsigned flags, float* state,
                          const float* torques, const float* ext_wrench, int N, int n_steps, float* out_foot_force,
                          void* cuda_stream) {
  if (!m) return fail(-1, "model handle is NULL");
  if (N <= 0 || n_steps < 0) return fail(-3, "N must be positive, n_steps non-negative");
  if (!state || !torques) return fail(-3, "state / torques is NULL");
  StepArgs A;
  std::memset(&A, 0, sizeof(A));
  if (int rc = make_ids(params ? P : 0, param_ids, &A.ids)) return rc;
  A.model = m->d_model; A.params = params; A.P = P; A.flags = flags;
  A.state = state; A.torques = torques; A.N = N; A.n_steps = n_steps; A.foot_force = out_foot_force;
  A.ext_wrench = ext_wrench;
  const int n_cta = (N + kRolloutsPerCta - 1) / kRolloutsPerCta;
  sim_step_kernel<<<n_cta, kThreads, 0, (cudaStream_t)cuda_stream>>>(A);
  return check_launch("sim_step_kernel");
}

int spi_b200_body_states(spi_b200_model* m, const float* state, int N, float* out, void* cuda_stream) {
  if (!m) return fail(-1, "model handle is NULL");
  if (N <= 0 || !state || !out) return fail(-3, "bad arguments");
  const int n = N * 5;
  body_states_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(m->fk, state, N, out);
  return check_launch("body_states_kernel");
}

int spi_b200_compute_torques(spi_b200_model* m, const float* actions, const float* q, const float* qd,
                             const float* gains, const float* motor_params, int N, int motor_model, unsigned flags,
                             float* out_tau, void* cuda_stream) {
  if (!m) return fail(-1, "model handle is NULL");
  if (N <= 0) return fail(-3, "N must be positive");
  if (!actions || !q || !qd || !out_tau) return fail(-3, "NULL buffer");
  if (motor_model < SPI_MOTOR_NONE || motor_model > SPI_MOTOR_VEC3_TANH) return fail(-3, "unknown motor_model");
  const int n = N * 12;
  torque_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(m->d_model, actions, q, qd, gains,
                                                                        motor_params, N, motor_model, flags, out_tau);
  return check_launch("torque_kernel");
}

int spi_b200_fim_reward(spi_b200_model* m, const float* states, int Mn, int P, float delta, int accumulate,
                        float* out_JtJ, float* out_trace, void* cuda_stream) {
  if (!m) return fail(-1, "model handle is NULL");
  if (Mn <= 0 || P <= 0) return fail(-3, "M and P must be positive");
  if (!states || (!out_JtJ && !out_trace)) return fail(-3, "NULL buffer");
  if (!(delta != 0.f)) return fail(-3, "delta must be non-zero");
  const int threads = 128, warps_per_cta = threads / 32;
  fim_reward_kernel<<<(Mn + warps_per_cta - 1) / warps_per_cta, threads, 0, (cudaStream_t)cuda_stream>>>(
      states, Mn, P, delta, accumulate, out_JtJ, out_trace);
  return check_launch("fim_reward_kernel");
}

int spi_b200_active_post_step(spi_b200_model* m, float* state, const float* raw_actions, unsigned char* done,
                              const float* main_commands, int T, float* commands, float* actions, float* gait,
                              float* clock, float* history, float* obs, void* obs_hi, void* obs_lo, int obs_stride,
                              int ring_slots, const int* hist_index, float* fim_hist,
                              unsigned char* fim_live, float* dead_steps, float* fim_jtj, float* fim_trace, float fim_delta,
                              const int* schedule, int schedule_rows, int* counter, int* ctrl,
                              int Mn, int P1, float dt, float action_clip, float clip_obs, float grav_x, float grav_y,
                              const float* q_default, void* cuda_stream) {
  if (!m) return fail(-1, "model handle is NULL");
  if (Mn <= 0 || P1 < 1 || P1 > activestep::kMaxGroup || T <= 0) return fail(-3, "bad M / group size / T");
  if (!state || !raw_actions || !done || !main_commands || !commands || !actions || !gait || !clock || !schedule ||
      !counter || !ctrl || !q_default)
    return fail(-3, "NULL buffer");
  if (ring_slots < 0 || (ring_slots > 0 && ring_slots != activestep::kHistLen + 1))
    return fail(-3, "ring_slots must be 0 or 15 (the frame + 14 history frames)");
  if (ring_slots > 0 && (!obs_hi || !obs_lo || obs_stride < ring_slots * activestep::kFrame))
    return fail(-3, "ring mode needs obs_hi / obs_lo with obs_stride >= 900");
  if (ring_slots == 0 && (!history || !obs || !hist_index)) return fail(-3, "NULL buffer");
  if (fim_hist && !fim_live) return fail(-3, "fim_live is NULL");
  if (fim_jtj && !(fim_delta != 0.f)) return fail(-3, "fim_delta must be non-zero with fim_jtj");
  if (obs_hi && (!obs_lo || obs_stride < activestep::kObs)) return fail(-3, "obs_lo is NULL or obs_stride < 900");
  static std::atomic<unsigned long long> attr_done{0};
  int attr_dev = -1;
  if (attr_needed_on_current_device(&attr_done, &attr_dev)) {
    CUDA_OK(cudaFuncSetAttribute(activestep::active_post_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)activestep::smem_bytes(activestep::kMaxGroup)));
    attr_mark_done(&attr_done, attr_dev);
  }
  activestep::Args A;
  A.state = state; A.raw_actions = raw_actions; A.done = done; A.main_commands = main_commands; A.commands = commands;
  A.actions = actions; A.gait = gait; A.clock = clock; A.history = history; A.obs = obs; A.hist_index = hist_index;
  A.obs_hi = static_cast<__half*>(obs_hi); A.obs_lo = static_cast<__half*>(obs_lo); A.obs_stride = obs_stride;
  A.obs_scale = mlptc::kActScale; A.ring_slots = ring_slots;
  A.fim_hist = fim_hist; A.fim_live = fim_live; A.dead_steps = dead_steps; A.ctrl = ctrl;
  A.schedule = schedule; A.schedule_rows = schedule_rows; A.counter = counter;
  A.fim_jtj = fim_jtj; A.fim_trace = fim_trace; A.fim_inv_delta = fim_delta != 0.f ? 1.0f / fim_delta : 0.f;
  A.M = Mn; A.P1 = P1; A.T = T; A.dt = dt; A.action_clip = action_clip; A.clip_obs = clip_obs;
  A.grav_x = grav_x; A.grav_y = grav_y;
  for (int j = 0; j < 12; j++) A.q_default[j] = q_default[j];
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CUDA_OK(pdl::launch(pdl_on(), activestep::active_post_step_kernel, dim3(Mn), dim3(32 * P1),
                      activestep::smem_bytes(P1, ring_slots > 0), st, A));
  return check_launch("active_post_step_kernel");
}

// ---- tensor-core policy MLP ---------------------------------------------------------------------------------------
struct spi_b200_policy {
  int dims[5] = {0, 0, 0, 0, 0};       // in, h1, h2, h3 (= 128), out
  int Kp = 0;                          // padded input width (multiple of 32)
  __half* w_hi[3] = {nullptr, nullptr, nullptr};   // x kWeightScale, fp16 pair, tiled
  __half* w_lo[3] = {nullptr, nullptr, nullptr};
  float* bias[3] = {nullptr, nullptr, nullptr};
  float* w_out = nullptr; float* b_out = nullptr;
  __half* act[4] = {nullptr, nullptr, nullptr, nullptr};  // h1 hi, h1 lo, h2 hi, h2 lo
  size_t act_rows = 0;
  std::vector<float> w1_host;          // layer-1 weights [h1, in] (ring copies are built from it)
  __half* ring_hi = nullptr; __half* ring_lo = nullptr; // [n_rot] tiled copies of layer 1 with permuted columns
  int n_rot = 0;
  size_t rot_stride = 0;               // elements between two copies
  // row chunks of a forward run as independent layer chains on side streams (policy_forward_impl)
  static constexpr int kMaxChunks = 8;
  cudaStream_t side[kMaxChunks - 1] = {};
  cudaEvent_t fork = nullptr, join[kMaxChunks - 1] = {};
  bool streams_ready = false;
};

static void policy_free(spi_b200_policy* p) {
  for (int l = 0; l < 3; l++) { cudaFree(p->w_hi[l]); cudaFree(p->w_lo[l]); cudaFree(p->bias[l]); }
  cudaFree(p->w_out); cudaFree(p->b_out);
  cudaFree(p->ring_hi); cudaFree(p->ring_lo);
  for (int i = 0; i < 4; i++) cudaFree(p->act[i]);
  if (p->streams_ready) {
    for (int i = 0; i < spi_b200_policy::kMaxChunks - 1; i++) { cudaStreamDestroy(p->side[i]); cudaEventDestroy(p->join[i]); }
    cudaEventDestroy(p->fork);
  }
  delete p;
}

static cudaError_t mlp_set_attributes() {
  cudaError_t e = cudaFuncSetAttribute(mlptc::mlp_layer_kernel<0, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, mlptc::kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mlptc::mlp_layer_kernel<0, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, mlptc::kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mlptc::mlp_layer_kernel<1, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, mlptc::kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mlptc::mlp_layer_kernel<0, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mlptc::kSmemBytes);
  return e;
}

// mode 0: hidden layer -> split activations (128 x 256 tiles when the width allows, else 128 x 128); mode 1: last hidden
// layer + output layer.  `layer` only labels the development timestamps.
static cudaError_t mlp_launch(int mode, int layer, mlptc::LayerArgs L, int m_tiles, cudaStream_t st) {
  static int wide = -1, pair = 0;
  if (wide < 0) {
    // SPI_B200_MLP_PAIR=1: CTA pairs (cta_group::2, 256 x 256 tiles per pair) for the wide layers.  Parity-tested, half the
    // operand bytes per flop and a 19 % shorter main loop per flop — but the 176 / 88 pair tiles of the config-5 batch quantise
    // worse over 148 SMs and clusters co-schedule less freely than single CTAs: 0.142 ms vs 0.108 ms per forward
    // (profiles/README.md r1_d), so it is off by default.
    const char* pe = std::getenv("SPI_B200_MLP_PAIR"); pair = pe ? std::atoi(pe) : 0;
    // 128 x 256 tiles: 25 % fewer operand bytes per flop but only 2 pipeline stages fit -> measured SLOWER (0.138 vs 0.112 ms)
    const char* w = std::getenv("SPI_B200_MLP_WIDE"); wide = w ? std::atoi(w) : 0;
  }
  L.dbg = 0; L.stamp = -1;
#if defined(SPI_B200_MLP_DEV)
  {
    static int dbg = -1, layers = 7, stamps = 0;
    if (dbg < 0) {
      const char* e = std::getenv("SPI_B200_MLP_DBG"); dbg = e ? std::atoi(e) : 0;
      const char* l = std::getenv("SPI_B200_MLP_LAYERS"); layers = l ? std::atoi(l) : 7;
      stamps = std::getenv("SPI_B200_MLP_STAMPS") ? 1 : 0;
    }
    L.dbg = dbg;
    if (!((layers >> layer) & 1)) return cudaSuccess;
    L.stamp = stamps ? layer : -1;
  }
#else
  (void)layer;
#endif
  if (mode == 1) {
    return pdl::launch(pdl_on(), mlptc::mlp_layer_kernel<1, 128>, dim3(m_tiles, 1), dim3(mlptc::kThreads), mlptc::kSmemBytes, st, L);
  } else if (pair && L.N % 256 == 0 && m_tiles % 2 == 0) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(m_tiles, L.N / 256); cfg.blockDim = dim3(mlptc::kThreads); cfg.dynamicSmemBytes = mlptc::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, mlptc::mlp_layer_kernel<0, 256, true>, L);
  } else if (wide && L.N % 256 == 0) {
    return pdl::launch(pdl_on(), mlptc::mlp_layer_kernel<0, 256>, dim3(m_tiles, L.N / 256), dim3(mlptc::kThreads), mlptc::kSmemBytes, st, L);
  } else {
    return pdl::launch(pdl_on(), mlptc::mlp_layer_kernel<0, 128>, dim3(m_tiles, L.N / 128), dim3(mlptc::kThreads), mlptc::kSmemBytes, st, L);
  }
  return cudaGetLastError();
}

// the fp16 pair of w * kWeightScale (host side of mlptc::split_half)
static void split_half_host(float w, __half* hi, __half* lo) {
  const float v = w * mlptc::kWeightScale;
  const __half h = __float2half_rn(v);
  *hi = h;
  *lo = __float2half_rn(v - __half2float(h));
}

int spi_b200_policy_create(const int* dims, const float* const* weights, const float* const* biases,
                           spi_b200_policy** out_policy) {
  if (!out_policy) return fail(-1, "out_policy is NULL");
  *out_policy = nullptr;
  if (!dims || !weights || !biases) return fail(-3, "NULL argument");
  for (int l = 0; l < 4; l++) if (!weights[l] || !biases[l]) return fail(-3, "NULL layer");
  if (dims[0] <= 0 || dims[1] <= 0 || dims[1] % mlptc::kTile || dims[2] <= 0 || dims[2] % mlptc::kTile ||
      dims[3] != mlptc::kTile || dims[4] <= 0 || dims[4] > mlptc::kMaxOut)
    return fail(-3, "tensor-core policy needs 3 hidden layers, widths h1, h2 multiples of 128, h3 = 128, <= 16 outputs");
  spi_b200_policy* p = new (std::nothrow) spi_b200_policy();
  if (!p) return fail(-4, "out of host memory");
  for (int i = 0; i < 5; i++) p->dims[i] = dims[i];
  // the fp16 pairs carry weights x 256: anything a trained actor holds fits, a weight beyond +-255 would overflow to inf
  for (int l = 0; l < 3; l++) {
    const size_t n = (size_t)dims[l + 1] * dims[l];
    for (size_t i = 0; i < n; i++)
      if (!(std::fabs(weights[l][i]) * mlptc::kWeightScale < 65504.0f)) {
        delete p;
        return fail(-3, "tensor-core policy: a hidden-layer weight is not finite or exceeds +-255 (fp16-pair range)");
      }
  }
  p->Kp = (dims[0] + mlptc::kKAlign - 1) / mlptc::kKAlign * mlptc::kKAlign;
  p->w1_host.assign(weights[0], weights[0] + (size_t)dims[1] * dims[0]);
  cudaError_t e = cudaSuccess;
  for (int l = 0; l < 3 && e == cudaSuccess; l++) {
    const int N = dims[l + 1], K = dims[l], Kp = (l == 0) ? p->Kp : K;
    std::vector<__half> hi((size_t)N * Kp, __float2half_rn(0.f)), lo((size_t)N * Kp, __float2half_rn(0.f));
    for (int n = 0; n < N; n++)
      for (int k = 0; k < K; k++) {
        const size_t o = tiled::offset(n, k, Kp);
        split_half_host(weights[l][(size_t)n * K + k], &hi[o], &lo[o]);
      }
    const size_t bytes = hi.size() * sizeof(__half);
    e = cudaMalloc((void**)&p->w_hi[l], bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->w_lo[l], bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->bias[l], (size_t)N * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(p->w_hi[l], hi.data(), bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->w_lo[l], lo.data(), bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->bias[l], biases[l], (size_t)N * sizeof(float), cudaMemcpyHostToDevice);
  }
  const size_t wo = (size_t)dims[4] * dims[3] * sizeof(float);
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->w_out, wo);
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->b_out, (size_t)dims[4] * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(p->w_out, weights[3], wo, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(p->b_out, biases[3], (size_t)dims[4] * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = mlp_set_attributes();
  if (e != cudaSuccess) {
    policy_free(p);
    return fail(-100, std::string("policy upload failed: ") + cudaGetErrorString(e));
  }
  *out_policy = p;
  return 0;
}

int spi_b200_policy_destroy(spi_b200_policy* p) {
  if (p) policy_free(p);
  return 0;
}

int spi_b200_policy_input_layout(spi_b200_policy* p, int M, int* out_rows, int* out_stride) {
  if (!p) return fail(-1, "policy handle is NULL");
  if (M <= 0) return fail(-3, "M must be positive");
  if (out_rows) *out_rows = (M + mlptc::kTile - 1) / mlptc::kTile * mlptc::kTile;
  if (out_stride) *out_stride = p->Kp;
  return 0;
}

int spi_b200_policy_split_input(spi_b200_policy* p, const float* x, int M, void* x_hi, void* x_lo, void* cuda_stream) {
  if (!p) return fail(-1, "policy handle is NULL");
  if (M <= 0 || !x || !x_hi || !x_lo) return fail(-3, "bad arguments");
  const size_t n = (size_t)M * p->dims[0];
  mlptc::split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(x, M, p->dims[0], static_cast<__half*>(x_hi), static_cast<__half*>(x_lo), p->Kp);
  return check_launch("split_kernel");
}

int spi_b200_policy_unsplit_input(spi_b200_policy* p, const void* x_hi, const void* x_lo, int M, float* x, void* cuda_stream) {
  if (!p) return fail(-1, "policy handle is NULL");
  if (M <= 0 || !x || !x_hi || !x_lo) return fail(-3, "bad arguments");
  const size_t n = (size_t)M * p->dims[0];
  mlptc::unsplit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(static_cast<const __half*>(x_hi), static_cast<const __half*>(x_lo), M, p->dims[0], p->Kp, x);
  return check_launch("unsplit_kernel");
}

int spi_b200_policy_enable_ring(spi_b200_policy* p, const int* col_map, int n_rot) {
  if (!p) return fail(-1, "policy handle is NULL");
  if (!col_map || n_rot < 1 || n_rot > 64) return fail(-3, "col_map is NULL or n_rot outside 1..64");
  const int N = p->dims[1], K = p->dims[0], Kp = p->Kp;
  for (size_t i = 0; i < (size_t)n_rot * K; i++)
    if (col_map[i] < -1 || col_map[i] >= K) return fail(-3, "col_map entry outside [-1, in)");
  const size_t stride = (size_t)N * Kp;
  std::vector<__half> hi(stride * n_rot, __float2half_rn(0.f)), lo(stride * n_rot, __float2half_rn(0.f));
  for (int r = 0; r < n_rot; r++)
    for (int n = 0; n < N; n++)
      for (int k = 0; k < K; k++) {
        const int c = col_map[(size_t)r * K + k];
        if (c < 0) continue;
        const size_t o = (size_t)r * stride + tiled::offset(n, k, Kp);
        split_half_host(p->w1_host[(size_t)n * K + c], &hi[o], &lo[o]);
      }
  cudaFree(p->ring_hi); cudaFree(p->ring_lo);
  p->ring_hi = p->ring_lo = nullptr; p->n_rot = 0;
  CUDA_OK(cudaMalloc((void**)&p->ring_hi, hi.size() * sizeof(__half)));
  CUDA_OK(cudaMalloc((void**)&p->ring_lo, lo.size() * sizeof(__half)));
  CUDA_OK(cudaMemcpy(p->ring_hi, hi.data(), hi.size() * sizeof(__half), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(p->ring_lo, lo.data(), lo.size() * sizeof(__half), cudaMemcpyHostToDevice));
  p->n_rot = n_rot; p->rot_stride = stride;
  return 0;
}

static int policy_forward_impl(spi_b200_policy* p, const void* x_hi, const void* x_lo, int M, bool ring,
                               const int* rot_dev, float* out, void* cuda_stream);

int spi_b200_policy_forward(spi_b200_policy* p, const void* x_hi, const void* x_lo, int M, float* out,
                            void* cuda_stream) {
  return policy_forward_impl(p, x_hi, x_lo, M, false, nullptr, out, cuda_stream);
}

int spi_b200_policy_forward_ring(spi_b200_policy* p, const void* x_hi, const void* x_lo, int M, const int* rot_dev,
                                 float* out, void* cuda_stream) {
  if (p && !p->n_rot) return fail(-3, "spi_b200_policy_enable_ring has not been called");
  return policy_forward_impl(p, x_hi, x_lo, M, true, rot_dev, out, cuda_stream);
}

static int policy_forward_impl(spi_b200_policy* p, const void* x_hi_v, const void* x_lo_v, int M, bool ring,
                               const int* rot_dev, float* out, void* cuda_stream) {
  const __half* x_hi = static_cast<const __half*>(x_hi_v);
  const __half* x_lo = static_cast<const __half*>(x_lo_v);
  if (!p) return fail(-1, "policy handle is NULL");
  if (M <= 0 || !x_hi || !x_lo || !out) return fail(-3, "bad arguments");
  const int Mp = (M + mlptc::kTile - 1) / mlptc::kTile * mlptc::kTile;
  if (p->act_rows < (size_t)Mp) {
    for (int i = 0; i < 4; i++) { cudaFree(p->act[i]); p->act[i] = nullptr; }
    p->act_rows = 0;
    for (int i = 0; i < 4; i++) CUDA_OK(cudaMalloc((void**)&p->act[i], (size_t)Mp * p->dims[1 + i / 2] * sizeof(__half)));
    p->act_rows = (size_t)Mp;
  }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  // The rows are cut into chunks of whole 128-row tiles and every chunk runs its own layer 1 -> 2 -> 3 chain on its own
  // stream: a layer's grid is 2.4 / 1.2 / 0.6 waves of one-CTA-per-SM tiles at the config-5 batch, and with a single chain
  // every partial wave idles most of the GPU; independent chains let the block scheduler fill those SMs with the next
  // layer of a chunk that is already done (measured: profiles/README.md).  Fork / join through events, so the whole
  // forward is still one dependency of `cuda_stream` and can be captured in a CUDA graph.
  const int m_tiles = Mp / mlptc::kTile;
  static int pref_chunks = 0;
  if (!pref_chunks) {
    const char* e = std::getenv("SPI_B200_MLP_CHUNKS");
    pref_chunks = e ? std::atoi(e) : 2;
    if (pref_chunks < 1) pref_chunks = 1;
    if (pref_chunks > spi_b200_policy::kMaxChunks) pref_chunks = spi_b200_policy::kMaxChunks;
  }
  const int n_chunks = (m_tiles >= 8 * pref_chunks) ? pref_chunks : 1;     // small batches: one chain
  if (n_chunks > 1 && !p->streams_ready) {
    for (int i = 0; i < spi_b200_policy::kMaxChunks - 1; i++) {
      CUDA_OK(cudaStreamCreateWithFlags(&p->side[i], cudaStreamNonBlocking));
      CUDA_OK(cudaEventCreateWithFlags(&p->join[i], cudaEventDisableTiming));
    }
    CUDA_OK(cudaEventCreateWithFlags(&p->fork, cudaEventDisableTiming));
    p->streams_ready = true;
  }
  if (n_chunks > 1) CUDA_OK(cudaEventRecord(p->fork, st));
  for (int c = 0; c < n_chunks; c++) {
    const int t0 = (int)((long long)m_tiles * c / n_chunks), t1 = (int)((long long)m_tiles * (c + 1) / n_chunks);
    const int tiles = t1 - t0, row0 = t0 * mlptc::kTile;
    if (tiles <= 0 || row0 >= M) continue;
    cudaStream_t cs = (c == 0) ? st : p->side[c - 1];
    if (c > 0) CUDA_OK(cudaStreamWaitEvent(cs, p->fork, 0));
    const size_t in_off = (size_t)t0 * (size_t)(p->Kp / 64) * tiled::kTileElems;
    const size_t h1_off = (size_t)t0 * (size_t)(p->dims[1] / 64) * tiled::kTileElems;
    const size_t h2_off = (size_t)t0 * (size_t)(p->dims[2] / 64) * tiled::kTileElems;
    mlptc::LayerArgs L;
    std::memset(&L, 0, sizeof(L));
    // layer 1
    L.a_hi = x_hi + in_off; L.a_lo = x_lo + in_off; L.w_hi = p->w_hi[0]; L.w_lo = p->w_lo[0]; L.bias = p->bias[0];
    L.Kp = p->Kp; L.N = p->dims[1];
    L.out_hi = p->act[0] + h1_off; L.out_lo = p->act[1] + h1_off; L.out_stride = p->dims[1]; L.M = M - row0;
    if (ring) { L.w_hi = p->ring_hi; L.w_lo = p->ring_lo; L.rot = rot_dev; L.rot_stride = p->rot_stride; L.n_rot = p->n_rot; }
    CUDA_OK(mlp_launch(0, 0, L, tiles, cs));
    if (int rc = check_launch("mlp_layer_kernel<0> (layer 1)")) return rc;
    // layer 2
    L.rot = nullptr; L.rot_stride = 0; L.n_rot = 0;
    L.a_hi = p->act[0] + h1_off; L.a_lo = p->act[1] + h1_off; L.w_hi = p->w_hi[1]; L.w_lo = p->w_lo[1]; L.bias = p->bias[1];
    L.Kp = p->dims[1]; L.N = p->dims[2];
    L.out_hi = p->act[2] + h2_off; L.out_lo = p->act[3] + h2_off; L.out_stride = p->dims[2];
    CUDA_OK(mlp_launch(0, 1, L, tiles, cs));
    if (int rc = check_launch("mlp_layer_kernel<0> (layer 2)")) return rc;
    // layer 3 + output layer
    L.a_hi = p->act[2] + h2_off; L.a_lo = p->act[3] + h2_off; L.w_hi = p->w_hi[2]; L.w_lo = p->w_lo[2]; L.bias = p->bias[2];
    L.Kp = p->dims[2]; L.N = p->dims[3];
    L.out_hi = nullptr; L.out_lo = nullptr; L.out_stride = 0;
    L.w_out = p->w_out; L.b_out = p->b_out; L.n_out = p->dims[4]; L.out = out + (size_t)row0 * p->dims[4];
    CUDA_OK(mlp_launch(1, 2, L, tiles, cs));
    if (int rc = check_launch("mlp_layer_kernel<1> (layers 3 + 4)")) return rc;
    if (c > 0) {
      CUDA_OK(cudaEventRecord(p->join[c - 1], cs));
      CUDA_OK(cudaStreamWaitEvent(st, p->join[c - 1], 0));
    }
  }
#if defined(SPI_B200_MLP_DEV)
  if (std::getenv("SPI_B200_MLP_STAMPS")) {
    unsigned long long h[3][8];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, mlptc::g_stamps, sizeof(h));
    for (int l = 0; l < 3; l++)
      std::fprintf(stderr, "[mlp stamps] layer %d: setup %llu  mainloop %llu  epilogue %llu (compute %llu)  exit %llu  | start-to-next-start %lld ns\n", l + 1,
                   h[l][1] - h[l][0], h[l][2] - h[l][1], h[l][3] - h[l][2], h[l][5] - h[l][2], h[l][4] - h[l][3], l < 2 ? (long long)(h[l + 1][0] - h[l][0]) : 0ll);
  }
#endif
  return check_launch("mlp_layer_kernel<1> (layers 3 + 4)");
}

int spi_b200_fim_contract(spi_b200_model* m, const float* hist, const unsigned char* live, int T, int Mn, int P,
                          float delta, int accumulate, float* out_JtJ, float* out_trace, void* cuda_stream) {
  if (!m) return fail(-1, "model handle is NULL");
  if (T <= 0 || Mn <= 0 || P <= 0) return fail(-3, "T, M and P must be positive");
  if (P > fimtc::kSlots) return fail(-3, "P must be <= 16");
  if (!hist || (!out_JtJ && !out_trace)) return fail(-3, "NULL buffer");
  if (!(delta != 0.f)) return fail(-3, "delta must be non-zero");
  static std::atomic<unsigned long long> attr_done{0};
  int attr_dev = -1;
  if (attr_needed_on_current_device(&attr_done, &attr_dev)) {
    CUDA_OK(cudaFuncSetAttribute(fimtc::fim_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 fimtc::kSmemBytes));
    attr_mark_done(&attr_done, attr_dev);
  }
  fimtc::FimArgs A;
  A.hist = hist; A.live = live; A.T = T; A.M = Mn; A.P = P; A.inv_delta = 1.0f / delta; A.accumulate = accumulate;
  A.out_JtJ = out_JtJ; A.out_trace = out_trace;
  const int n_cta = (Mn + fimtc::kEnvsPerCta - 1) / fimtc::kEnvsPerCta;
  // slices of the control steps: ONE wave of 2 CTAs per SM (measured at 128 env tiles, T = 1 248: 2 slices 1.00 ms, 3 slices
  // 1.30 ms, 5 slices 1.18 ms, 8 slices 1.01 ms — a second, partial wave costs more than the finer slices win), >= 4 steps each
  static const int split_env = [] { const char* e = getenv("SPI_B200_FIM_SPLIT"); return e ? atoi(e) : 0; }();
  int n_split = split_env > 0 ? split_env : (2 * m->sm_count) / n_cta;
  { const int cap = T / 4 > 1 ? T / 4 : 1; if (n_split > cap) n_split = cap; if (n_split < 1) n_split = 1; }
  A.n_split = n_split; A.part_JtJ = nullptr; A.part_trace = nullptr;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const size_t n_jtj = (size_t)Mn * P * P;
  if (n_split > 1) {
    spi_b200_model::FimPart* fp;
    { std::lock_guard<std::mutex> lock(m->fim_part_mu); fp = &m->fim_part[st]; }
    if (int rc = ensure(&fp->ptr, &fp->cap, (size_t)n_split * (n_jtj + Mn))) return rc;
    A.part_JtJ = fp->ptr; A.part_trace = fp->ptr + (size_t)n_split * n_jtj;
  }
  fimtc::fim_contract_kernel<<<dim3(n_cta, n_split), fimtc::kRows, fimtc::kSmemBytes, st>>>(A);
  if (int rc = check_launch("fim_contract_kernel")) return rc;
  if (n_split > 1) {
    if (out_JtJ) fimtc::fim_reduce_kernel<<<(unsigned)((n_jtj + 255) / 256), 256, 0, st>>>(A.part_JtJ, n_split, n_jtj, accumulate, out_JtJ);
    if (out_trace) fimtc::fim_reduce_kernel<<<(Mn + 255) / 256, 256, 0, st>>>(A.part_trace, n_split, (size_t)Mn, accumulate, out_trace);
    return check_launch("fim_reduce_kernel");
  }
  return 0;
}

int spi_b200_weighted_cost(spi_b200_model* m, const float* cost3, int C, float w_pos, float w_quat, float w_joint,
                           float* out_total, void* cuda_stream) {
  if (!m) return fail(-1, "model handle is NULL");
  if (C <= 0 || !cost3 || !out_total) return fail(-3, "bad arguments");
  weighted_cost_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(cost3, C, w_pos, w_quat, w_joint,
                                                                               out_total);
  return check_launch("weighted_cost_kernel");
}

int spi_b200_cem_sample(spi_b200_model* m, const float* mean, const float* std, const float* lo, const float* hi, int C,
                        int P, int c0, unsigned long long seed, int iteration, float* out_params, void* cuda_stream) {
  if (!m) return fail(-1, "model handle is NULL");
  if (C <= 0 || P <= 0 || !mean || !std || !out_params) return fail(-3, "bad arguments");
  const int n = C * P;
  cem_sample_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(mean, std, lo, hi, C, P, c0, seed,
                                                                            iteration, out_params);
  return check_launch("cem_sample_kernel");
}

int spi_b200_cem_refit(spi_b200_model* m, const float* params, const float* cost, int C, int P, int n_elite,
                       float alpha, const float* std_floor, float* mean, float* std, float* out_best,
                       void* cuda_stream) {
  if (!m) return fail(-1, "model handle is NULL");
  if (C <= 0 || P <= 0 || n_elite <= 0 || n_elite > C) return fail(-3, "bad C / P / n_elite");
  if (!params || !cost || !mean || !std) return fail(-3, "NULL buffer");
  if (int rc = ensure(&m->d_rank, &m->rank_cap, (size_t)C)) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  cem_select_kernel<<<1, kSelThreads, 0, st>>>(cost, C, n_elite, m->d_rank);
  if (int rc = check_launch("cem_select_kernel")) return rc;
  cem_refit_kernel<<<P, 256, 0, st>>>(params, cost, m->d_rank, C, P, n_elite, alpha, std_floor, mean, std, out_best);
  return check_launch("cem_refit_kernel");
}

int spi_b200_fp32_peak(int iters, float* out_tflops, float* out_ms, void* cuda_stream) {
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int dev = 0, sms = 0;
  CUDA_OK(cudaGetDevice(&dev));
  CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float* sink = nullptr;
  CUDA_OK(cudaMalloc((void**)&sink, 4));
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  const int grid = sms * 8, threads = 256;
  fp32_peak_kernel<<<grid, threads, 0, st>>>(iters / 4 + 1, 1.0f, sink);  // warm-up
  g_launches.fetch_add(1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CUDA_OK(cudaEventRecord(e0, st));
    fp32_peak_kernel<<<grid, threads, 0, st>>>(iters, 1.0f, sink);
    g_launches.fetch_add(1);
    CUDA_OK(cudaEventRecord(e1, st));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
  const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)grid * (double)threads;
  if (out_tflops) *out_tflops = (float)(flops / (best * 1e-3) / 1e12);
  if (out_ms) *out_ms = best;
  return 0;
}

}  // extern "C"

#if defined(SPI_WS_PROFILE)
// dev builds only (not part of include/spi_b200.h): device buffer [nblk][5][256] of clock64 stamps for CTAs blk0 .. blk0 + nblk
extern "C" int spi_b200_debug_ws_prof(long long* buf, int blk0, int nblk) {
  g_ws_prof = buf; g_ws_prof_blk0 = blk0; g_ws_prof_nblk = nblk;
  return 0;
}
#endif
