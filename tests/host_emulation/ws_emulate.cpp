// CPU emulation of the warp-specialised rollout kernel (TEST INFRASTRUCTURE): runs the very same
// __host__ __device__ role functions of spi_active_b200/csrc/go2_ws.cuh for ONE rollout, role by role, with
// plain arrays in place of shared memory, in the order the kernel's barriers impose.  Built with g++ by
// tests/test_ws_emulation.py and compared against the oracle.
#include <cstring>

#include "../../spi_active_b200/csrc/go2_ws.cuh"

using namespace ws;

extern "C" int ws_emulate_rollout(const float* blob, const float* params, int P, const int* ids, unsigned flags,
                                  int motor_model, const float* init, const float* actions, const float* gains, int H,
                                  int decimation, float* out_states, float* out_foot_force) {
  ModelK M;
  std::memset(&M, 0, sizeof(M));
  if (int rc = model_from_blob(blob, &M)) return rc;
  ParamIdsK pid;
  pid.n = P;
  for (int i = 0; i < 16; i++) pid.id[i] = i < P ? ids[i] : -1;
  BaseInertia B;
  float motor[3];
  apply_candidate(M, P > 0 ? params : nullptr, pid, flags, B, motor);
  if (motor_model == SPI_MOTOR_SCALAR) { motor[1] = motor[0]; motor[2] = motor[0]; }
  BaseState bs;
  for (int i = 0; i < 3; i++) { bs.p[i] = init[i]; bs.v[i] = init[7 + i]; bs.w[i] = init[10 + i]; }
  for (int i = 0; i < 4; i++) bs.quat[i] = init[3 + i];
  LegState ls[4];
  float kp[4][3], kd[4][3];
  for (int leg = 0; leg < 4; leg++)
    for (int j = 0; j < 3; j++) {
      ls[leg].q[j] = init[13 + 3 * leg + j];
      ls[leg].qd[j] = init[25 + 3 * leg + j];
      kp[leg][j] = gains ? gains[3 * leg + j] : M.kp[3 * leg + j];
      kd[leg][j] = gains ? gains[12 + 3 * leg + j] : M.kd[3 * leg + j];
    }
  float bc[kBaseOut];
  for (int i = 0; i < kBaseOut; i++) bc[i] = 0.f;
  base_publish(bs, bc);
  float pb[6];
  base_bias(B, bc, pb);
  const float h = M.sim.dt / (float)M.sim.nsub;
  LegKeep K[4];
  float part[4][kLegOut], ff[4][3];
  for (int k = 0; k < H; k++) {
    float act[4][3];
    for (int leg = 0; leg < 4; leg++)
      for (int j = 0; j < 3; j++)
        act[leg][j] = fminf(fmaxf(actions[12 * k + 3 * leg + j], -M.sim.action_clip), M.sim.action_clip);
    for (int d = 0; d < decimation; d++) {
      float tau[4][3];
      for (int leg = 0; leg < 4; leg++)
        leg_torques(M.sim, M.leg[leg], act[leg], ls[leg].q, ls[leg].qd, kp[leg], kd[leg], motor, motor_model, flags, tau[leg]);
      for (int n = 0; n < M.sim.nsub; n++) {
        for (int leg = 0; leg < 4; leg++) leg_phase1(M.sim, M.leg[leg], bc, ls[leg], tau[leg], K[leg], part[leg], ff[leg]);
        float legsum[kLegOut];
        for (int i = 0; i < kLegOut; i++) legsum[i] = (part[0][i] + part[1][i]) + (part[2][i] + part[3][i]);
        base_solve(B, legsum, pb, bc + kBcA0);      // between barriers A and B1
        float a0[6];
        for (int i = 0; i < 6; i++) a0[i] = bc[kBcA0 + i];
        base_advance(M.sim, a0, bs, h, bc);         // between barriers B1 and B2 (legs run phase 2 meanwhile)
        base_bias(B, bc, pb);
        for (int leg = 0; leg < 4; leg++) leg_phase2(M.leg[leg], bc, K[leg], ls[leg], h);
      }
    }
    float* o = out_states + 37 * k;
    for (int i = 0; i < 3; i++) { o[i] = bs.p[i]; o[7 + i] = bs.v[i]; o[10 + i] = bs.w[i]; }
    for (int i = 0; i < 4; i++) o[3 + i] = bs.quat[i];
    for (int leg = 0; leg < 4; leg++)
      for (int j = 0; j < 3; j++) { o[13 + 3 * leg + j] = ls[leg].q[j]; o[25 + 3 * leg + j] = ls[leg].qd[j]; }
  }
  if (out_foot_force)
    for (int leg = 0; leg < 4; leg++)
      for (int i = 0; i < 3; i++) out_foot_force[3 * leg + i] = ff[leg][i];
  return 0;
}
