// active_step.cuh — everything LeggedRobotBase / go2_omni / ActiveSysId_OpenLoop do after the physics of one control
// step of the active-exploration rollout (BASELINE config 5), fused into one kernel.  The reference runs ~130 eager
// torch kernels per step for this bookkeeping; here it is one launch, one CTA per (1 main + P aux) env group, one warp
// per env:
//
//   _update_tasks_callback      commands <- command row of the new step      active_sysid_openloop.py:196-201
//   _check_termination          |projected gravity x / y| > 0.8 (+ non-finite state), OR-ed over the group
//                                                                            legged_robot_base.py:336-339, active_sysid_openloop.py:259-272
//   FIM inputs                  (root13, q12) of every env -> history ring for spi_b200_fim_contract (:402-426)
//   _compute_observations       60 scaled terms in SORTED key order, 14-frame newest-first history gather, clip +-100
//                                                                            legged_robot_base.py:240-250, 511-527, 819-829;
//                                                                            config/obs/loco/go2_omni.yaml; utils/helpers.py:77-94
//   history_handler.add         push the (unclipped, scaled) frame           env_utils/history_handler.py:36-44
//   k-step sync                 aux <- main every k steps (intent of :247-252, 316-330; quirk D11)
//   _step_contact_targets       gait clock of the NEXT step, from the new commands   go2_omni.py:348-377
//
// The arithmetic follows spi_active_b200/active.py (the torch statement of the same step, pinned to the reference's
// own code by tests/golden/active_obs.npz) operation by operation, so the two paths agree to fp32 rounding.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "tiled_layout.cuh"
#include "pdl.cuh"

namespace activestep {

constexpr int kFrame = 60, kHistLen = 14, kObs = kFrame * (1 + kHistLen);   // 900
constexpr int kState = 37, kFimDim = 25;
constexpr int kMaxGroup = 17;   // 1 main + <= 16 aux envs (the tensor-core contraction's slot count)

struct Args {
  float* state;                    // [N,37] in/out (sync)
  const float* raw_actions;        // [N,12] policy output of this step
  unsigned char* done;             // [N] in: flags of the previous step (whose envs had their action zeroed), out: new
  const float* main_commands;      // [M,T,14] command trajectories of the main envs
  float* commands;                 // [N,14] out: command row in force after this step
  float* actions;                  // [N,12] out: the clipped / zeroed action the physics used
  float* gait;                     // [N] in/out
  float* clock;                    // [N,4] in/out
  float* history;                  // [N,14,60] in/out, newest first
  float* obs;                      // [N,900] out
  __half* obs_hi; __half* obs_lo;  // [rows, obs_stride] (tiled_layout.cuh) or null: pre-split input of spi_b200_policy_forward
  float obs_scale;                 // power of two applied before the fp16 split (mlptc::kActScale)
  int obs_stride;
  // ring_slots > 0 (= 15): obs_hi / obs_lo ARE the observation state — a ring of 15 frames per env, K position
  // slot * 60 + term; this step's frame goes to slot ctrl[3] and nothing else is touched (history / obs / hist_index
  // are not used).  The actor's first layer then runs with the weight columns permuted for that head position
  // (spi_b200_policy_enable_ring): the 900-dim observation [frame | per-key history blocks] is never materialised.
  int ring_slots;
  const int* hist_index;           // [840] gather index of short_history into the flattened [14*60] ring
  float* fim_hist;                 // [K,M,P1,25] or null
  unsigned char* fim_live;         // [K,M] or null
  float* dead_steps;               // [N] or null
  // fused Fisher accumulation (alternative to fim_hist + spi_b200_fim_contract): jtj[m] += live * J J^T of THIS step, trace[m] +=
  // its trace, with J = (main - aux_p) * inv_delta in R^{P x 25} (active_sysid_openloop.py:402-426) formed from the state rows
  // the kernel already holds in shared memory — no [T, M, P1, 25] history is written or read back
  float* fim_jtj;                  // [M,P,P] or null
  float* fim_trace;                // [M] or null
  float fim_inv_delta;
  // The per-step host inputs come from a schedule uploaded once per rollout: this step uses row counter[0] (the last row once
  // the counter runs past it).  The LAST block to finish (ticket counter[1]) publishes that row in ctrl — [0] command row t,
  // [1] k-sync flag, [2] FIM ring slot, [3] observation ring head: what the actor's first layer reads next — and advances
  // counter[0]: no separate one-thread kernel between the physics and this one
  const int* schedule; int schedule_rows;
  int* counter;                    // [2]: step counter, block ticket (0 between launches)
  int* ctrl;                       // [4] out
  int M, P1, T;
  float dt, action_clip, clip_obs, grav_x, grav_y;
  float q_default[12];
};

__device__ __forceinline__ float remainder1(float x) {   // torch.remainder(x, 1.0)
  float r = fmodf(x, 1.0f);
  if (r != 0.f && r < 0.f) r += 1.0f;
  return r;
}

// spigym/utils/torch_utils.py:83-92 (q = xyzw)
__device__ __forceinline__ void quat_rotate_inverse(const float* q, const float* v, float* o) {
  const float w = q[3];
  const float s = 2.0f * (w * w) - 1.0f;
  const float cx = q[1] * v[2] - q[2] * v[1], cy = q[2] * v[0] - q[0] * v[2], cz = q[0] * v[1] - q[1] * v[0];
  const float dot = (q[0] * v[0] + q[1] * v[1]) + q[2] * v[2];
  o[0] = (v[0] * s - cx * w * 2.0f) + q[0] * dot * 2.0f;
  o[1] = (v[1] * s - cy * w * 2.0f) + q[1] * dot * 2.0f;
  o[2] = (v[2] * s - cz * w * 2.0f) + q[2] * dot * 2.0f;
}

// per-env shared-memory slice: state row (40) | frame (64) | history (840; not in ring mode), + one flag per env at the end
constexpr int kEnvSmemFloats = 40 + 64 + kHistLen * kFrame, kEnvSmemFloatsRing = 40 + 64;
inline size_t smem_bytes(int P1, bool ring = false) {
  return (size_t)P1 * ((ring ? kEnvSmemFloatsRing : kEnvSmemFloats) * sizeof(float) + sizeof(int));
}

// (the per-step host inputs — command row, k-sync flag, FIM ring slot, ring head — come from a schedule uploaded once per rollout: Args)
__global__ void __launch_bounds__(32 * kMaxGroup) active_post_step_kernel(const Args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl::trigger();
  pdl::wait();
  const bool ring = A.ring_slots > 0;
  struct View {
    float* base; int* flag; int stride;
    __device__ float* st(int w) const { return base + (size_t)w * stride; }
    __device__ float* frame(int w) const { return st(w) + 40; }
    __device__ float* hist(int w) const { return st(w) + 104; }
  } sm;
  sm.base = reinterpret_cast<float*>(smem_raw);
  sm.stride = ring ? kEnvSmemFloatsRing : kEnvSmemFloats;
  sm.flag = reinterpret_cast<int*>(sm.base + (size_t)A.P1 * sm.stride);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x;
  const int env = m * A.P1 + w;
  // (counter[0] only changes when the last block of this launch is done, i.e. after every block has read it)
  const int c0 = A.counter[0];
  const int* sched = A.schedule + 4 * (c0 < A.schedule_rows ? c0 : A.schedule_rows - 1);
  const int t_cmd = sched[0], do_sync = sched[1], slot = sched[2], ring_head = sched[3];

  // ---- state row, termination ------------------------------------------------------------------------------------
  float* srow = A.state + (size_t)env * kState;
  bool finite = true;
  for (int i = lane; i < kState; i += 32) {
    const float v = srow[i];
    sm.st(w)[i] = v;
    finite = finite && (fabsf(v) <= 3.0e38f);   // false for NaN and Inf
  }
  finite = __all_sync(0xffffffffu, finite);
  __syncwarp();
  const float* s = sm.st(w);
  const float gdown[3] = {0.f, 0.f, -1.0f};
  float pg[3], ang[3];
  quat_rotate_inverse(s + 3, gdown, pg);
  quat_rotate_inverse(s + 3, s + 10, ang);
  if (lane == 0) sm.flag[w] = (fabsf(pg[0]) > A.grav_x || fabsf(pg[1]) > A.grav_y || !finite) ? 1 : 0;
  const bool was_done = A.done[env] != 0;
  __syncthreads();
  bool group_done = false;
  for (int i = 0; i < A.P1; i++) group_done = group_done || (sm.flag[i] != 0);

  // ---- FIM inputs ----------------------------------------------------------------------------------------------------
  if (A.fim_hist) {
    if (lane < kFimDim) A.fim_hist[(((size_t)slot * A.M + m) * A.P1 + w) * kFimDim + lane] = s[lane];
    if (w == 0 && lane == 0) A.fim_live[(size_t)slot * A.M + m] = group_done ? 0 : 1;
  }
  if ((A.fim_hist || A.fim_jtj) && lane == 0 && A.dead_steps) A.dead_steps[env] += group_done ? 1.0f : 0.0f;
  if (A.fim_jtj && !group_done) {
    // one (p, q) entry per thread: 25-term dot product of two difference rows (the rows sit in shared memory: sm.st(i)[0 .. 25))
    const int P = A.P1 - 1;
    const float* main_row = sm.st(0);
    for (int idx = threadIdx.x; idx < P * P; idx += blockDim.x) {
      const int p = idx / P, q = idx - p * P;
      const float* ap = sm.st(p + 1);
      const float* aq = sm.st(q + 1);
      float acc = 0.f;
#pragma unroll 5
      for (int d = 0; d < kFimDim; d++)
        acc = fmaf((main_row[d] - ap[d]) * A.fim_inv_delta, (main_row[d] - aq[d]) * A.fim_inv_delta, acc);
      A.fim_jtj[((size_t)m * P + p) * P + q] += acc;
    }
    if (w == 0 && A.fim_trace) {             // trace = ||J||_F^2, summed in a fixed order (lane p, then a shuffle tree)
      float t = 0.f;
      if (lane < P) {
        const float* ap = sm.st(lane + 1);
#pragma unroll 5
        for (int d = 0; d < kFimDim; d++) { const float j = (main_row[d] - ap[d]) * A.fim_inv_delta; t = fmaf(j, j, t); }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) A.fim_trace[m] += t;
    }
  }

  // ---- the 60 scaled observation terms ---------------------------------------------------------------------------------
  const float* cmd = A.main_commands + ((size_t)m * A.T + t_cmd) * 14;
  const float g = A.gait[env];
  float height = cmd[3] + 0.20f * sinf(6.2831855f * g);
  height = fminf(fmaxf(height, -0.25f), 0.15f);
  for (int i = lane; i < kFrame; i += 32) {
    float v;
    if (i < 12) {
      const float a = A.raw_actions[(size_t)env * 12 + i];
      // a non-finite raw action (the reference's fp32 actor cannot produce one from finite, clipped observations) acts as 0
      v = (was_done || !(fabsf(a) <= 3.0e38f)) ? 0.f : fminf(fmaxf(a, -A.action_clip), A.action_clip);
      A.actions[(size_t)env * 12 + i] = v;
    } else if (i < 15) v = ang[i - 12] * 0.25f;
    else if (i < 19) v = A.clock[(size_t)env * 4 + (i - 15)];
    else if (i == 19) v = cmd[2];
    else if (i < 22) v = cmd[10 + (i - 20)] * 0.3f;
    else if (i == 22) v = height * 2.0f;
    else if (i == 23) v = cmd[9] * 0.15f;
    else if (i == 24) v = cmd[4];
    else if (i < 29) v = cmd[5 + (i - 25)];
    else if (i < 31) v = cmd[i - 29];
    else if (i < 33) v = cmd[12 + (i - 31)];
    else if (i < 45) v = s[13 + (i - 33)] - A.q_default[i - 33];
    else if (i < 57) v = s[25 + (i - 45)] * 0.05f;
    else v = pg[i - 57];
    sm.frame(w)[i] = v;
  }
  if (lane < 14) A.commands[(size_t)env * 14 + lane] = cmd[lane];

  if (ring) {
    // ---- ring mode: the clipped frame, split, into slot ctrl[3] of the actor's input operand -------------------------------
    __syncwarp();
    const int k0 = ring_head * kFrame;
    for (int i = lane; i < kFrame; i += 32) {
      const float o = fminf(fmaxf(sm.frame(w)[i], -A.clip_obs), A.clip_obs) * A.obs_scale;
      const __half h = __float2half_rn(o);
      const size_t at = tiled::offset(env, k0 + i, A.obs_stride);
      A.obs_hi[at] = h;
      A.obs_lo[at] = __float2half_rn(o - __half2float(h));
    }
  } else {
  // ---- observation = [frame | gathered history], clip; then push the frame -----------------------------------------------
  float* hrow = A.history + (size_t)env * (kHistLen * kFrame);
  for (int i = lane; i < kHistLen * kFrame; i += 32) sm.hist(w)[i] = hrow[i];
  __syncwarp();
  float* orow = A.obs + (size_t)env * kObs;
  for (int i = lane; i < kObs; i += 32) {
    const float v = (i < kFrame) ? sm.frame(w)[i] : sm.hist(w)[A.hist_index[i - kFrame]];
    const float o = fminf(fmaxf(v, -A.clip_obs), A.clip_obs);
    orow[i] = o;
    if (A.obs_hi) {
      const float os = o * A.obs_scale;
      const __half h = __float2half_rn(os);
      const size_t at = tiled::offset(env, i, A.obs_stride);
      A.obs_hi[at] = h;
      A.obs_lo[at] = __float2half_rn(os - __half2float(h));
    }
  }
  for (int i = lane; i < kHistLen * kFrame; i += 32) hrow[i] = (i < kFrame) ? sm.frame(w)[i] : sm.hist(w)[i - kFrame];
  }

  // ---- termination flag, k-step sync, gait clock of the next step --------------------------------------------------------
  if (lane == 0) A.done[env] = group_done ? 1 : 0;
  if (do_sync && w > 0)
    for (int i = lane; i < kState; i += 32) srow[i] = sm.st(0)[i];
  if (lane == 0) {
    const float freq = cmd[4], phases = cmd[5], offsets = cmd[6], bounds = cmd[7], dur = cmd[8];
    const float gn = remainder1(g + A.dt * freq);
    const float foot[4] = {((gn + phases) + offsets) + bounds, gn + offsets, gn + bounds, gn + phases};
    A.gait[env] = gn;
#pragma unroll
    for (int f = 0; f < 4; f++) {
      const float r = remainder1(foot[f]);
      float warped = foot[f];
      if (r < dur) warped = r * (0.5f / dur);
      if (r > dur) warped = 0.5f + (r - dur) * (0.5f / (1.0f - dur));
      A.clock[(size_t)env * 4 + f] = sinf(6.2831855f * warped);
    }
  }

  // ---- the last block publishes the step's schedule row and advances the counter ------------------------------------------------
  // (no fence: the ticket only has to order every block's READ of counter[0] / the schedule row — complete before its threads
  // reach the barrier, since they used the values — before the last block's WRITE; a __threadfence() here also invalidates the
  // SM's L1 under the blocks that are still running: +10 us on the 1 024-block launch)
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(A.counter + 1, 1);
    if (ticket == (int)gridDim.x - 1) {
      A.ctrl[0] = t_cmd; A.ctrl[1] = do_sync; A.ctrl[2] = slot; A.ctrl[3] = ring_head;
      A.counter[0] = c0 < A.schedule_rows ? c0 + 1 : c0;
      A.counter[1] = 0;
    }
  }
}

}  // namespace activestep
