// oracle/spi_oracle_capi.cpp — C entry points of the CPU oracle (TEST INFRASTRUCTURE, see
// spi_oracle.hpp).  Built by oracle/Makefile into oracle/_build/libspi_oracle.so and loaded with
// ctypes from oracle/oracle.py.  Mirrors the argument lists of include/spi_b200.h with HOST pointers.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <vector>
#include <atomic>
#include <thread>

#include "spi_oracle.hpp"

using namespace spi_oracle;

namespace {

// minimal parallel-for over [0,total) in chunks, std::thread based (no OpenMP dependency)
template <class Fn> void parallel_for(long long total, int n_threads, long long chunk, Fn fn) {
  if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
  if (n_threads < 1) n_threads = 1;
  if ((long long)n_threads * chunk > total) n_threads = (int)std::max<long long>(1, total / std::max<long long>(1, chunk));
  if (n_threads <= 1) { for (long long i = 0; i < total; i++) fn(i); return; }
  std::atomic<long long> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; t++)
    pool.emplace_back([&]() {
      for (;;) {
        long long b = next.fetch_add(chunk);
        if (b >= total) break;
        long long e = std::min(total, b + chunk);
        for (long long i = b; i < e; i++) fn(i);
      }
    });
  for (auto& th : pool) th.join();
}

// scripts/eval.py:279-310 — masked sums over the segments divided by `total_valid`
template <class T>
int eval_candidates_t(const float* blob, int n_floats, const float* params, int C, int P, const int* ids,
                      const float* seg_init, const float* seg_actions, const float* seg_target,
                      const float* seg_gains, const unsigned char* seg_mask, int S, int H, int decimation,
                      int motor_model, unsigned flags, float cost_denominator, double* out_cost,
                      double* out_per_seg, int* out_status, int n_threads) {
  Model<T> M;
  if (!M.load(blob, n_floats)) return -1;
  double denom = cost_denominator;
  if (!(denom > 0.0)) {
    denom = 0.0;
    for (int s = 0; s < S; s++) denom += seg_mask ? (seg_mask[s] ? 1.0 : 0.0) : 1.0;
  }
  std::vector<double> per((size_t)C * S * 3);
  std::vector<int> bad((size_t)C, 0);
  std::vector<std::atomic<int>> bad_flags((size_t)C);
  for (auto& b : bad_flags) b.store(0);
  const long long total = (long long)C * S;
  parallel_for(total, n_threads, 64, [&](long long idx) {
    int c = (int)(idx / S), s = (int)(idx % S);
    Candidate<T> cand = apply_params(M, params + (size_t)c * P, P, ids, flags);
    State<T> st;
    rollout<T, float>(M, cand, seg_init + (size_t)s * SPI_STATE_DIM, seg_actions + (size_t)s * H * 12,
                      seg_gains ? seg_gains + (size_t)s * 24 : nullptr, H, decimation, motor_model, flags, st,
                      nullptr);
    T err[3];
    segment_errors(st, seg_target + (size_t)s * SPI_TARGET_DIM, err);
    bool ok = st.finite();
    for (int k = 0; k < 3; k++) per[((size_t)c * S + s) * 3 + k] = to_double(err[k]);
    if (!ok) bad_flags[c].store(1);
  });
  for (int c = 0; c < C; c++) bad[c] = bad_flags[c].load();
  for (int c = 0; c < C; c++) {
    double sum[3] = {0, 0, 0};
    for (int s = 0; s < S; s++) {
      double m = seg_mask ? (seg_mask[s] ? 1.0 : 0.0) : 1.0;
      if (m != 0.0)
        for (int k = 0; k < 3; k++) sum[k] += per[((size_t)c * S + s) * 3 + k];
    }
    for (int k = 0; k < 3; k++)
      out_cost[c * 3 + k] = bad[c] ? std::numeric_limits<double>::infinity() : sum[k] / denom;
    if (out_status) out_status[c] = bad[c];
  }
  if (out_per_seg) std::copy(per.begin(), per.end(), out_per_seg);
  return 0;
}

template <class T>
int rollout_states_t(const float* blob, int n_floats, const float* params, int C, int P, const int* ids,
                     const float* seg_init, const float* seg_actions, const float* seg_gains, int S, int H,
                     int decimation, int motor_model, unsigned flags, double* out_states) {
  Model<T> M;
  if (!M.load(blob, n_floats)) return -1;
  const long long total = (long long)C * S;
  parallel_for(total, 0, 16, [&](long long idx) {
    int c = (int)(idx / S), s = (int)(idx % S);
    Candidate<T> cand = apply_params(M, params + (size_t)c * P, P, ids, flags);
    State<T> st;
    rollout<T, double>(M, cand, seg_init + (size_t)s * SPI_STATE_DIM, seg_actions + (size_t)s * H * 12,
                       seg_gains ? seg_gains + (size_t)s * 24 : nullptr, H, decimation, motor_model, flags, st,
                       out_states + (size_t)idx * H * SPI_STATE_DIM);
  });
  return 0;
}

}  // namespace

extern "C" {

int spi_oracle_eval_candidates(int precision /*64|32*/, const float* blob, int n_floats, const float* params, int C,
                               int P, const int* ids, const float* seg_init, const float* seg_actions,
                               const float* seg_target, const float* seg_gains, const unsigned char* seg_mask,
                               int S, int H, int decimation, int motor_model, unsigned flags,
                               float cost_denominator, double* out_cost, double* out_per_seg, int* out_status,
                               int n_threads) {
  if (precision == 32)
    return eval_candidates_t<float>(blob, n_floats, params, C, P, ids, seg_init, seg_actions, seg_target, seg_gains,
                                    seg_mask, S, H, decimation, motor_model, flags, cost_denominator, out_cost,
                                    out_per_seg, out_status, n_threads);
  return eval_candidates_t<double>(blob, n_floats, params, C, P, ids, seg_init, seg_actions, seg_target, seg_gains,
                                   seg_mask, S, H, decimation, motor_model, flags, cost_denominator, out_cost,
                                   out_per_seg, out_status, n_threads);
}

int spi_oracle_rollout_states(int precision, const float* blob, int n_floats, const float* params, int C, int P,
                              const int* ids, const float* seg_init, const float* seg_actions,
                              const float* seg_gains, int S, int H, int decimation, int motor_model, unsigned flags,
                              double* out_states) {
  if (precision == 32)
    return rollout_states_t<float>(blob, n_floats, params, C, P, ids, seg_init, seg_actions, seg_gains, S, H,
                                   decimation, motor_model, flags, out_states);
  return rollout_states_t<double>(blob, n_floats, params, C, P, ids, seg_init, seg_actions, seg_gains, S, H,
                                  decimation, motor_model, flags, out_states);
}

// stepwise simulator: N envs, n_steps physics steps under constant torques (double precision)
int spi_oracle_sim_step_ext(const float* blob, int n_floats, const float* params, int P, const int* ids, unsigned flags,
                            double* state, const double* torques, int N, int n_steps, double* out_foot_force,
                            const double* ext_wrench /*[N,13,6] body-frame [torque; force] per moving body, or null*/);
int spi_oracle_sim_step(const float* blob, int n_floats, const float* params, int P, const int* ids, unsigned flags,
                        double* state /*[N,37] in/out*/, const double* torques /*[N,12]*/, int N, int n_steps,
                        double* out_foot_force /*[N,4,3] or null*/) {
  return spi_oracle_sim_step_ext(blob, n_floats, params, P, ids, flags, state, torques, N, n_steps, out_foot_force, nullptr);
}
int spi_oracle_sim_step_ext(const float* blob, int n_floats, const float* params, int P, const int* ids, unsigned flags,
                            double* state, const double* torques, int N, int n_steps, double* out_foot_force,
                            const double* ext_wrench) {
  Model<double> M;
  if (!M.load(blob, n_floats)) return -1;
  parallel_for(N, 0, 16, [&](long long e) {
    Candidate<double> cand = apply_params(M, params ? params + (size_t)e * P : nullptr, params ? P : 0, ids, flags);
    Bodies<double> Ib = make_bodies(M, cand);
    State<double> st;
    double* sp = state + (size_t)e * SPI_STATE_DIM;
    st.p = V3<double>(sp[0], sp[1], sp[2]);
    for (int k = 0; k < 4; k++) st.quat[k] = sp[3 + k];
    st.v = V3<double>(sp[7], sp[8], sp[9]);
    st.w = V3<double>(sp[10], sp[11], sp[12]);
    for (int j = 0; j < 12; j++) { st.q[j] = sp[13 + j]; st.qd[j] = sp[25 + j]; }
    V3<double> ff[4];
    SV<double> ext[13];
    if (ext_wrench)
      for (int b = 0; b < 13; b++) {
        const double* w = ext_wrench + ((size_t)e * 13 + b) * 6;
        ext[b].a = V3<double>(w[0], w[1], w[2]); ext[b].l = V3<double>(w[3], w[4], w[5]);
      }
    for (int k = 0; k < n_steps; k++) physics_step(M, Ib, st, torques + (size_t)e * 12, ff, ext_wrench ? ext : nullptr);
    st.to(sp);
    if (out_foot_force)
      for (int l = 0; l < 4; l++) {
        out_foot_force[((size_t)e * 4 + l) * 3 + 0] = ff[l].x;
        out_foot_force[((size_t)e * 4 + l) * 3 + 1] = ff[l].y;
        out_foot_force[((size_t)e * 4 + l) * 3 + 2] = ff[l].z;
      }
  });
  return 0;
}

// one CONTROL step of N independent envs, each with its own parameter row (closed-loop / active-exploration path):
// LeggedRobotBase.step's control path (legged_robot_base.py:185-209) in double precision, state in/out
int spi_oracle_env_step(const float* blob, int n_floats, const float* params, int P, const int* ids, double* state,
                        const float* actions, const float* gains, int N, int decimation, int motor_model,
                        unsigned flags) {
  Model<double> M;
  if (!M.load(blob, n_floats)) return -1;
  parallel_for(N, 0, 8, [&](long long e) {
    Candidate<double> cand = apply_params(M, params ? params + (size_t)e * P : nullptr, params ? P : 0, ids, flags);
    Bodies<double> Ib = make_bodies(M, cand);
    State<double> st;
    double* sp = state + (size_t)e * SPI_STATE_DIM;
    st.p = V3<double>(sp[0], sp[1], sp[2]);
    for (int k = 0; k < 4; k++) st.quat[k] = sp[3 + k];
    st.v = V3<double>(sp[7], sp[8], sp[9]);
    st.w = V3<double>(sp[10], sp[11], sp[12]);
    for (int j = 0; j < 12; j++) { st.q[j] = sp[13 + j]; st.qd[j] = sp[25 + j]; }
    double kp[12], kd[12], a[12], tau[12];
    for (int j = 0; j < 12; j++) {
      kp[j] = gains ? (double)gains[(size_t)e * 24 + j] : M.kp[j];
      kd[j] = gains ? (double)gains[(size_t)e * 24 + 12 + j] : M.kd[j];
      a[j] = s_clip((double)actions[(size_t)e * 12 + j], -M.action_clip, M.action_clip);
    }
    for (int d = 0; d < decimation; d++) {
      compute_torques(M, a, st.q, st.qd, kp, kd, cand.motor, motor_model, flags, tau);
      physics_step(M, Ib, st, tau);
    }
    st.to(sp);
  });
  return 0;
}

// forward dynamics of one state (for the CRBA/RNEA cross-check and the invariants tests):
// out_acc = base spatial acceleration in base coords [ang3, lin3] (gravity included if requested),
// out_qdd[12], out_foot[4,3] world contact forces
int spi_oracle_forward_dynamics(const float* blob, int n_floats, const float* params, int P, const int* ids,
                                unsigned flags, const double* state, const double* tau, int with_contact,
                                int with_gravity, double* out_acc, double* out_qdd, double* out_foot) {
  Model<double> M;
  if (!M.load(blob, n_floats)) return -1;
  Candidate<double> cand = apply_params(M, params, P, ids, flags);
  Bodies<double> Ib = make_bodies(M, cand);
  State<double> st;
  st.p = V3<double>(state[0], state[1], state[2]);
  for (int k = 0; k < 4; k++) st.quat[k] = state[3 + k];
  st.v = V3<double>(state[7], state[8], state[9]);
  st.w = V3<double>(state[10], state[11], state[12]);
  for (int j = 0; j < 12; j++) { st.q[j] = state[13 + j]; st.qd[j] = state[25 + j]; }
  Accel<double> A = forward_dynamics(M, Ib, st, tau, with_contact != 0, with_gravity != 0);
  for (int k = 0; k < 3; k++) { out_acc[k] = A.a0.a[k]; out_acc[3 + k] = A.a0.l[k]; }
  for (int j = 0; j < 12; j++) out_qdd[j] = A.qdd[j];
  if (out_foot)
    for (int l = 0; l < 4; l++)
      for (int k = 0; k < 3; k++) out_foot[l * 3 + k] = A.foot_force[l][k];
  return 0;
}

// torque law on its own: N rows
int spi_oracle_compute_torques(int precision, const float* blob, int n_floats, const float* actions, const float* q,
                               const float* qd, const float* gains, const float* motor_params, int N,
                               int motor_model, unsigned flags, double* out_tau) {
  Model<double> Md; Model<float> Mf;
  if (!Md.load(blob, n_floats) || !Mf.load(blob, n_floats)) return -1;
  for (int e = 0; e < N; e++) {
    if (precision == 32) {
      float a[12], qq[12], qqd[12], kp[12], kd[12], mot[3], tau[12];
      for (int j = 0; j < 12; j++) {
        a[j] = s_clip(actions[e * 12 + j], -Mf.action_clip, Mf.action_clip);
        qq[j] = q[e * 12 + j]; qqd[j] = qd[e * 12 + j];
        kp[j] = gains ? gains[e * 24 + j] : Mf.kp[j]; kd[j] = gains ? gains[e * 24 + 12 + j] : Mf.kd[j];
      }
      for (int k = 0; k < 3; k++) mot[k] = motor_params ? motor_params[e * 3 + k] : 20.0f;
      compute_torques(Mf, a, qq, qqd, kp, kd, mot, motor_model, flags, tau);
      for (int j = 0; j < 12; j++) out_tau[e * 12 + j] = tau[j];
    } else {
      double a[12], qq[12], qqd[12], kp[12], kd[12], mot[3], tau[12];
      for (int j = 0; j < 12; j++) {
        a[j] = s_clip((double)actions[e * 12 + j], -Md.action_clip, Md.action_clip);
        qq[j] = q[e * 12 + j]; qqd[j] = qd[e * 12 + j];
        kp[j] = gains ? gains[e * 24 + j] : Md.kp[j]; kd[j] = gains ? gains[e * 24 + 12 + j] : Md.kd[j];
      }
      for (int k = 0; k < 3; k++) mot[k] = motor_params ? motor_params[e * 3 + k] : 20.0;
      compute_torques(Md, a, qq, qqd, kp, kd, mot, motor_model, flags, tau);
      for (int j = 0; j < 12; j++) out_tau[e * 12 + j] = tau[j];
    }
  }
  return 0;
}

// algorithmic FLOPs of ONE (candidate, segment) rollout, counted by running the scalar code on the
// op-counting type.  out_counts = {add, mul, div, sqrt, transcendental, compare}
int spi_oracle_count_flops(const float* blob, int n_floats, const float* params, int P, const int* ids,
                           const float* init, const float* actions, const float* target, int H, int decimation,
                           int motor_model, unsigned flags, unsigned long long* out_counts) {
  Model<Counted> M;
  if (!M.load(blob, n_floats)) return -1;
  op_counts() = OpCounts();
  Candidate<Counted> cand = apply_params(M, params, P, ids, flags);
  State<Counted> st;
  rollout<Counted, float>(M, cand, init, actions, nullptr, H, decimation, motor_model, flags, st, nullptr);
  Counted err[3];
  segment_errors(st, target, err);
  OpCounts c = op_counts();
  out_counts[0] = c.add; out_counts[1] = c.mul; out_counts[2] = c.div;
  out_counts[3] = c.sqrt; out_counts[4] = c.trans; out_counts[5] = c.cmp;
  return 0;
}

int spi_oracle_num_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
