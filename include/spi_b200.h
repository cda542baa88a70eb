/*
 * spi_b200.h — C-ABI of the B200-native SysID rollout engine (libspi_b200.so).
 *
 * This is the drop-in boundary "B2" of SURVEY.md §8(b): the fused candidate-evaluation
 * operator that replaces, in the reference (LeCAR-Lab/SPI-Active, paths relative to its root):
 *
 *   scripts/eval.py:204-214   apply_base_mass          -> a row of `params` per candidate
 *   scripts/eval.py:217-310   evaluate_batch           -> spi_b200_eval_candidates
 *   scripts/mass_landscape.py:111-128  mass_sweep      -> one call with C = 20 candidates
 *   scripts/mass_opt.py:136-169 evaluate_mass_scale    -> one call with C = 1 (or a batch of trials)
 *   spigym/envs/legged_base_task/legged_robot_base.py:169-209,529-560
 *                         step / _physics_step / _compute_torques   -> fused in the rollout kernel
 *   spigym/envs/sysid/active_sysid_openloop.py:174-187,356-400
 *                         motor models act2tau_*       -> `motor_model` enum, fused
 *   spigym/envs/sysid/active_sysid_openloop.py:402-426
 *                         _reward_fisher_information_matrix -> spi_b200_fim_* (J^T J on tensor cores)
 *   spigym/simulator/isaacgym/isaacgym.py:598-626
 *                         apply_torques_at_dof / simulate_at_each_physics_step -> spi_b200_sim_step
 *   spigym/simulator/isaacgym/isaacgym_active_sysid.py:61-94
 *                         set_mass / set_com* / set_inertia*  -> `param_ids`
 *
 * Conventions
 *   - plain C, no torch types.  All `const float*` / `float*` data arguments of the device
 *     entry points are DEVICE pointers on the current CUDA device, caller-owned, fp32,
 *     row-major, 16-byte aligned.  `param_ids` and `model_blob` are HOST pointers.
 *   - calls are asynchronous on `cuda_stream` (a cudaStream_t passed as void*); nothing is
 *     allocated after spi_b200_model_create except when a call needs a larger workspace than
 *     any before it (grown once, never shrunk).
 *   - return 0 on success, negative on error; message via spi_b200_last_error() (thread-local).
 *   - one stream per model handle at a time.
 *   - the `*_host` entry points take HOST pointers and do the H2D/D2H copies themselves
 *     (pinned staging owned by the handle); they synchronise the stream before returning.
 */
#ifndef SPI_B200_H
#define SPI_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SPI_B200_VERSION 105 /* major*100 + minor */

/* ------------------------------------------------------------------------------------------
 * Model blob layout (fp32[SPI_BLOB_SIZE]).  Built on the host from the URDF
 * (spigym/data/robots/go2/urdf/go2.urdf) and spigym/config/robot/go2/go2.yaml:29-128 by
 * spi_active_b200.go2_model.build_model_blob().
 * 13 moving bodies: base + 4 legs (FL, FR, RL, RR) x (hip, thigh, calf); feet are lumped into
 * the calves at blob-build time, the two head links are lumped into the base at run time
 * (after the candidate's base-link parameters are applied).
 * Inertial record = mass, com[3], I[6] = (xx, yy, zz, xy, xz, yz) about the com, link axes.
 * ------------------------------------------------------------------------------------------ */
enum {
  SPI_BLOB_MAGIC = 0,          /* = 20025.0f */
  SPI_BLOB_DT = 1,             /* physics step, s (0.005) */
  SPI_BLOB_GRAVITY_Z = 2,      /* -9.81 */
  SPI_BLOB_ACTION_SCALE = 3,   /* 0.25 */
  SPI_BLOB_ACTION_CLIP = 4,    /* 20.0 */
  SPI_BLOB_CONTACT_KN = 5,     /* normal stiffness N/m */
  SPI_BLOB_CONTACT_CN = 6,     /* Hunt-Crossley damping factor s/m: f_n = kn*d*(1 - cn*vz) */
  SPI_BLOB_CONTACT_MU = 7,     /* Coulomb friction coefficient */
  SPI_BLOB_CONTACT_DT = 8,     /* tangential viscous coefficient N s/m (capped by mu*f_n) */
  SPI_BLOB_FOOT_RADIUS = 9,    /* 0.022 */
  SPI_BLOB_NSUB = 10,          /* integrator sub-steps per physics step (float-encoded int) */
  SPI_BLOB_CONTACT_VEPS = 11,  /* slip-speed regulariser m/s */
  SPI_BLOB_FOOT_SPHERE = 12,   /* 3 floats: foot collision-sphere centre in the FOOT link frame (urdf: -0.002, 0, 0) */
  SPI_BLOB_BASE_INERTIAL = 16, /* 10 floats: URDF base link */
  SPI_BLOB_BASE_LUMPS = 26,    /* 2 x 10 floats: mass, pos[3] in base frame, I[6] about own com */
  SPI_BLOB_LEG_BODIES = 46,    /* 12 x 14 floats: inertial[10], joint origin in parent[3], axis id (0=x,1=y) */
  SPI_BLOB_FOOT_OFFSET = 214,  /* 4 x 3: foot sphere centre in calf frame */
  SPI_BLOB_Q_DEFAULT = 226,    /* 12 */
  SPI_BLOB_TORQUE_LIMIT = 238, /* 12 */
  SPI_BLOB_KP = 250,           /* 12 */
  SPI_BLOB_KD = 262,           /* 12 */
  SPI_BLOB_Q_LOWER = 274,      /* 12 */
  SPI_BLOB_Q_UPPER = 286,      /* 12 */
  SPI_BLOB_QD_LIMIT = 298,     /* 12 */
  SPI_BLOB_SIZE = 312
};
#define SPI_BLOB_MAGIC_VALUE 20025.0f
#define SPI_LEG_BODY_STRIDE 14
#define SPI_INERTIAL_STRIDE 10

/* Candidate parameter ids (columns of `params`).  Names follow
 * spigym/config/env/active_sysid_openloop.yaml:17-27 and the setters in
 * spigym/simulator/isaacgym/isaacgym_active_sysid.py:61-94. */
enum {
  SPI_PARAM_MASS = 0,          /* base-link mass, kg                       (set_mass)     */
  SPI_PARAM_COMX = 1,          /* base-link centre of mass, m              (set_comx)     */
  SPI_PARAM_COMY = 2,
  SPI_PARAM_COMZ = 3,
  SPI_PARAM_INERTIAX = 4,      /* base-link Ixx about com, kg m^2          (set_inertiax) */
  SPI_PARAM_INERTIAY = 5,      /*   (the reference's `set_inertiaiy` typo drops this one: D8) */
  SPI_PARAM_INERTIAZ = 6,
  SPI_PARAM_INERTIAXY = 7,     /* off-diagonals: beyond the reference, default to the URDF */
  SPI_PARAM_INERTIAXZ = 8,
  SPI_PARAM_INERTIAYZ = 9,
  SPI_PARAM_MOTOR_HIP = 10,    /* motor_model_hip_a   | hip_gain   | scalar_gain */
  SPI_PARAM_MOTOR_THIGH = 11,  /* motor_model_thigh_a | thigh_gain */
  SPI_PARAM_MOTOR_CALF = 12,   /* motor_model_calf_a  | calf_gain  */
  SPI_PARAM_MASS_SCALE = 13,   /* base-link mass as a multiple of the URDF mass (mass_opt's `mass_scale`) */
  SPI_PARAM_COUNT = 14
};

/* motor models: spigym/envs/sysid/active_sysid_openloop.py:356-400 */
enum {
  SPI_MOTOR_NONE = 0,        /* plain LeggedRobotBase: PD -> clip            (scripts/eval.py path) */
  SPI_MOTOR_SCALAR = 1,      /* act2tau_scalar:    tau * g                   (g = SPI_PARAM_MOTOR_HIP) */
  SPI_MOTOR_VEC3 = 2,        /* act2tau_vec3:      tau * g_{hip,thigh,calf} */
  SPI_MOTOR_VEC3_TANH = 3    /* act2tau_vec3_tanh: a * tanh(tau / a) applied after the clip */
};

/* flags */
enum {
  SPI_FLAG_HIP_HALF = 1u << 0,          /* go2_omni.py:436-437: hip actions x0.5 */
  SPI_FLAG_INERTIA_KEEP = 1u << 1,      /* D15: keep the URDF inertia when only the mass changes
                                           (default: scale inertia with mass)                  */
  SPI_FLAG_STRICT_INERTIAY = 1u << 2,   /* D8: ignore SPI_PARAM_INERTIAY like the reference    */
  SPI_FLAG_TANH_BEFORE_CLIP = 1u << 3   /* go2_locomotion.py:108-111 order: PD -> tanh -> clip */
};

/* state row layout used by rollout_states / sim_step: 37 floats per env
 *   pos[3] quat_xyzw[4] lin_vel_world[3] ang_vel_world[3] q[12] qd[12]
 * (root-state row of isaacgym.py:567-572 followed by dof pos / vel)                          */
#define SPI_STATE_DIM 37
#define SPI_TARGET_DIM 19 /* pos[3] quat[4] q[12] */
#define SPI_NQ 12

typedef struct spi_b200_model spi_b200_model; /* opaque */

int spi_b200_version(void);
const char* spi_b200_last_error(void);

/* host blob -> device-resident model + workspace.  Replaces asset loading / env creation:
 * spigym/simulator/isaacgym/isaacgym.py:170-272 (load_assets, create_envs). */
int spi_b200_model_create(const float* model_blob, int n_floats, spi_b200_model** out_model);
int spi_b200_model_destroy(spi_b200_model* model);

/* Fused hot path.  For every candidate c in [0,C) and segment s in [0,S): reset to seg_init[s],
 * replay seg_actions[s, 0..H) with `decimation` physics steps per action, compare to
 * seg_target[s]; masked mean over s.  Replaces scripts/eval.py:217-310 + :204-214 +
 * scripts/mass_landscape.py:123-126.
 *   params      [C,P] device; param_ids [P] HOST (values SPI_PARAM_*)
 *   seg_init    [S,37] device; seg_actions [S,H,12]; seg_target [S,19];
 *   seg_gains   [S,24] (kp[12], kd[12]) or NULL -> blob defaults
 *   seg_mask    [S] bytes, 1 = sample counts (eval_mask of eval.py:279-280), or NULL -> all 1
 *   cost_denominator : divisor of the masked sums (eval.py:304 `total_valid`); <= 0 -> sum(mask)
 *   out_cost    [C,3] mean (base_pos, base_quat, joint_pos) L2 errors
 *   out_per_seg [C,S,3] unmasked per-sample errors, or NULL
 *   out_status  [C] int: 0 ok, 1 = some rollout went non-finite (cost = +inf), or NULL        */
int spi_b200_eval_candidates(spi_b200_model* model,
                             const float* params, int C, int P, const int* param_ids,
                             const float* seg_init, const float* seg_actions,
                             const float* seg_target, const float* seg_gains,
                             const unsigned char* seg_mask, int S, int H, int decimation,
                             int motor_model, unsigned flags, float cost_denominator,
                             float* out_cost, float* out_per_seg, int* out_status,
                             void* cuda_stream);

/* Same call with HOST pointers everywhere (the reference-facing plugin call: numpy in,
 * numpy out).  H2D of params + dataset and D2H of the costs happen inside. */
int spi_b200_eval_candidates_host(spi_b200_model* model,
                                  const float* params, int C, int P, const int* param_ids,
                                  const float* seg_init, const float* seg_actions,
                                  const float* seg_target, const float* seg_gains,
                                  const unsigned char* seg_mask, int S, int H, int decimation,
                                  int motor_model, unsigned flags, float cost_denominator,
                                  float* out_cost, int* out_status, void* cuda_stream);

/* Per-step parity hook / dataset recorder: state after each of the H control steps.
 *   out_states [C,S,H,37]                                                                     */
int spi_b200_rollout_states(spi_b200_model* model,
                            const float* params, int C, int P, const int* param_ids,
                            const float* seg_init, const float* seg_actions,
                            const float* seg_gains, int S, int H, int decimation,
                            int motor_model, unsigned flags,
                            float* out_states, void* cuda_stream);

/* Stepwise simulator boundary "B1" (BaseSimulator.apply_torques_at_dof +
 * simulate_at_each_physics_step, spigym/simulator/base_simulator/base_simulator.py:121,150):
 * advance N envs by `n_steps` physics steps under constant joint torques.
 *   params [N,P] per-env rigid-body parameters (or NULL -> URDF), state [N,37] in/out,
 *   torques [N,12], out_foot_force [N,4,3] world-frame contact force of the last step or NULL */
int spi_b200_sim_step(spi_b200_model* model,
                      const float* params, int P, const int* param_ids, unsigned flags,
                      float* state, const float* torques, int N, int n_steps,
                      float* out_foot_force, void* cuda_stream);

/* The same step with external wrenches (IsaacGym.apply_rigid_body_force_at_pos_tensor,
 * spigym/simulator/isaacgym/isaacgym.py:609-613): ext_wrench [N,13,6] or NULL = [torque; force] on each of the 13 moving
 * bodies (base, then hip / thigh / calf of FL, FR, RL, RR; forces on feet / head links belong to the calf / base they are
 * fixed to) expressed in the body's own link frame about its link origin, held constant over the n_steps physics steps.
 * spi_active_b200.simulator.B200Sim converts the reference's world-frame (force, position) pairs of the 19 Isaac Gym bodies. */
int spi_b200_sim_step_ext(spi_b200_model* model,
                          const float* params, int P, const int* param_ids, unsigned flags,
                          float* state, const float* torques, const float* ext_wrench, int N, int n_steps,
                          float* out_foot_force, void* cuda_stream);

/* Rigid-body state tensor of the 19 Isaac Gym bodies (spigym/simulator/isaacgym/isaacgym.py:541-577
 * `_rigid_body_pos / _rot / _vel / _ang_vel`; body order of spigym/config/robot/go2/go2.yaml:44: base, FL hip / thigh /
 * calf / foot, FR ..., Head_upper, Head_lower, RL ..., RR ...) by forward kinematics of the engine state.
 *   state [N,37] -> out [N,19,13] = link-frame origin pos[3], link-frame quat_xyzw[4], linear velocity OF THE LINK
 *   ORIGIN [3], angular velocity [3], all in the world frame.                                                       */
int spi_b200_body_states(spi_b200_model* model, const float* state, int N, float* out, void* cuda_stream);

/* One CONTROL step of N independent envs, each with its own parameter row (the closed-loop / active-exploration path:
 * LeggedRobotBase.step without the observation / reward bookkeeping, legged_robot_base.py:169-209, with the torque law
 * of the env in use, go2_omni.py:423-465 + active_sysid_openloop.py:174-187): clip the action, then `decimation` x
 * (PD + motor model from the fresh q, qd -> one physics step).
 *   params [N,P] or NULL, state [N,37] in/out, actions [N,12], gains [N,24] or NULL -> blob defaults;
 *   zero_action_mask [N] bytes or NULL: 1 = the env's action is replaced by 0 before the clip (the driver zeroes the
 *   actions of terminated envs, agents/sysid/active_sysid.py:559-562)                                              */
int spi_b200_env_step(spi_b200_model* model, const float* params, int P, const int* param_ids, float* state,
                      const float* actions, const unsigned char* zero_action_mask, const float* gains, int N,
                      int decimation, int motor_model, unsigned flags, void* cuda_stream);

/* Motor-model + PD torque operator on its own (parity hook for
 * legged_robot_base.py:545,557 and active_sysid_openloop.py:356-400):
 *   tau[N,12] = motor(clip(kp*(scale*a + q_default - q) - kd*qd))                             */
int spi_b200_compute_torques(spi_b200_model* model,
                             const float* actions, const float* q, const float* qd,
                             const float* gains /* [N,24] or NULL */,
                             const float* motor_params /* [N,3] or NULL */,
                             int N, int motor_model, unsigned flags,
                             float* out_tau, void* cuda_stream);

/* Fisher-information reward of active exploration
 * (active_sysid_openloop.py:402-426): states [M,(P+1),25] = root13 (origin-compensated) + q12
 * for main + P aux envs; J = (main - aux)/delta in R^{P x 25};
 *   out_JtJ [M,P,P] (+= if accumulate != 0) and out_trace [M] = ||J||_F^2 (the reward).       */
int spi_b200_fim_reward(spi_b200_model* model, const float* states, int M, int P, float delta,
                        int accumulate, float* out_JtJ, float* out_trace, void* cuda_stream);

/* Everything the env classes do AFTER the physics of one control step of the active-exploration rollout, fused:
 * command playback (active_sysid_openloop.py:196-201), gravity termination OR-ed over each (1 main + P aux) group
 * (legged_robot_base.py:336-339, active_sysid_openloop.py:259-272), the FIM inputs of the step (:402-426) into the
 * history ring of spi_b200_fim_contract, the 900-dim actor observation (legged_robot_base.py:240-250, 511-527, 819-829,
 * config/obs/loco/go2_omni.yaml) + history push (env_utils/history_handler.py:36-44), the k-step aux <- main sync
 * (:247-252, 316-330) and the gait clock of the next step (go2_omni.py:348-377).  N = M * P1 envs, group-major.
 *   state [N,37] in/out; raw_actions [N,12] = policy output the physics of this step used; done [N] bytes in: flags of
 *   the previous step (those envs ran with a zero action), out: new flags; main_commands [M,T,14];
 *   commands [N,14], actions [N,12] out; gait [N], clock [N,4], history [N,14,60] in/out; obs [N,900] out;
 *   obs_hi / obs_lo [rows, obs_stride] or NULL: the same observation pre-split for spi_b200_policy_forward;
 *   ring_slots: 0, or 15 = RING MODE: obs_hi / obs_lo are the observation state itself, a ring of 15 clipped frames per env
 *   (element slot * 60 + term); the step writes its frame into slot ctrl[3] and touches nothing else — history, obs and
 *   hist_index may be NULL — and the actor runs through spi_b200_policy_forward_ring with the matching head position;
 *   hist_index [840] device ints (short_history gather); fim_hist [K,M,P1,25] + fim_live [K,M] or NULL;
 *   dead_steps [N] or NULL (+= done); schedule [steps,4] device ints (command row, sync flag, FIM slot, ring head),
 *   fim_jtj [M,P,P] / fim_trace [M] (or NULL): FUSED Fisher accumulation — jtj[m] += live * J J^T of this step, trace[m] += its
 *   trace, J = (main - aux_p) / fim_delta in R^{P x 25} (active_sysid_openloop.py:402-426), from the rows the kernel already
 *   holds; the alternative to recording fim_hist / fim_live for spi_b200_fim_contract (pass those as NULL then);
 *   counter [2] / ctrl [4] device ints: the step uses row schedule[counter[0]]; the last block of the kernel publishes that row
 *   in ctrl (ctrl[3] = the ring head spi_b200_policy_forward_ring reads next) and advances counter[0]; counter[1] is the block
 *   ticket of that hand-over and must be 0 at launch (it is 0 again afterwards).  schedule has `schedule_rows` rows of 4 ints —
 *   a step past the last row re-reads the last row instead of running off the buffer;
 *   q_default [12] HOST.  1 <= P1 <= 17.                                                                          */
int spi_b200_active_post_step(spi_b200_model* model, float* state, const float* raw_actions, unsigned char* done,
                              const float* main_commands, int T, float* commands, float* actions, float* gait,
                              float* clock, float* history, float* obs, void* obs_hi, void* obs_lo, int obs_stride,
                              int ring_slots, const int* hist_index, float* fim_hist,
                              unsigned char* fim_live, float* dead_steps, float* fim_jtj, float* fim_trace, float fim_delta,
                              const int* schedule, int schedule_rows, int* counter, int* ctrl, int M, int P1, float dt, float action_clip, float clip_obs, float grav_x, float grav_y,
                              const float* q_default, void* cuda_stream);

/* The locomotion policy (actor MLP: Linear-ELU x3 + Linear; spigym/agents/modules/modules.py:47-63,
 * config/algo/ppo.yaml:32-40, evaluated per control step at agents/sysid/active_sysid.py:555-562) on the tensor cores
 * with fp32-grade accuracy: tcgen05.mma kind::f16 on fp16 PAIRS (x s = hi + lo, products hi.hi + hi.lo + lo.hi accumulated in
 * fp32: ~2^-22 relative, where plain fp16 / TF32 would be 2^-11).  Constraints: 3 hidden layers, widths h1, h2 multiples of
 * 128, h3 = 128, <= 16 outputs (the reference's 900-512-256-128-12 actor qualifies); |weights| < 255 and |activations| < 4000
 * (the pairs are scaled by 256 / 16 to keep the low halves out of the fp16 subnormals; observations are clipped to +-100).
 *   dims [5] = in, h1, h2, h3, out; weights[l] [dims[l+1], dims[l]] row-major (torch Linear layout), biases[l]: HOST.
 * The input is passed PRE-SPLIT in two OPAQUE buffers of rows * stride 2-byte elements each
 * (spi_b200_policy_input_layout: rows = M rounded up to 128, stride = in rounded up to 64; zero-initialised by the owner):
 * 128 x 64 tiles stored as the swizzled shared-memory image the tensor core reads, so that a pipeline stage is four
 * contiguous 16 KB TMA bulk copies (csrc/tiled_layout.cuh).  spi_b200_active_post_step writes the observation in that form
 * directly; spi_b200_policy_split_input / _unsplit_input convert from / to a plain row-major [M, in] fp32 matrix.
 * out [M, out] fp32.                                                                                              */
typedef struct spi_b200_policy spi_b200_policy;
int spi_b200_policy_create(const int* dims, const float* const* weights, const float* const* biases,
                           spi_b200_policy** out_policy);
int spi_b200_policy_destroy(spi_b200_policy* policy);
int spi_b200_policy_input_layout(spi_b200_policy* policy, int M, int* out_rows, int* out_stride);
int spi_b200_policy_split_input(spi_b200_policy* policy, const float* x, int M, void* x_hi, void* x_lo,
                                void* cuda_stream);
int spi_b200_policy_unsplit_input(spi_b200_policy* policy, const void* x_hi, const void* x_lo, int M, float* x,
                                  void* cuda_stream);
int spi_b200_policy_forward(spi_b200_policy* policy, const void* x_hi, const void* x_lo, int M, float* out,
                            void* cuda_stream);

/* Ring-ordered input (active exploration): the actor's input [frame | per-key blocks of the 14 previous frames] is a
 * fixed permutation of a ring of the last 15 frames, so instead of rebuilding the 900-dim observation every control step
 * (legged_robot_base.py:511-527, 819-829: ~17 KB of memory traffic per env and step) the caller keeps the ring as the
 * pre-split operand, writes one frame per step, and the first layer's weight columns are permuted instead — one copy per
 * head position.  col_map [n_rot][dims[0]] HOST: col_map[r][k] = the original input column whose weight multiplies ring
 * element k when the head is at position r (-1 = unused element).  rot_dev: DEVICE int holding the head position of
 * this forward (read by the kernel, so a captured step needs no host work), or NULL for position 0.                  */
int spi_b200_policy_enable_ring(spi_b200_policy* policy, const int* col_map, int n_rot);
int spi_b200_policy_forward_ring(spi_b200_policy* policy, const void* x_hi, const void* x_lo, int M,
                                 const int* rot_dev, float* out, void* cuda_stream);

/* Accumulated Fisher information of whole rollouts on the tensor cores (tcgen05.mma kind::tf32, 3xTF32 split =
 * fp32-accurate): the sum over control steps of the per-step J J^T that active_sysid_openloop.py:402-426 forms and
 * active_sysid.py:567-590 accumulates (terminated groups contribute termination_rew = 0 -> `live`).
 *   hist [T,M,(P+1),25] the per-step `states` of spi_b200_fim_reward stacked over T control steps;
 *   live [T,M] bytes (1 = the group's step counts) or NULL -> all 1;  P <= 16;
 *   out_JtJ [M,P,P] = sum_t live * J_t J_t^T (+= if accumulate != 0), out_trace [M] = its trace (the summed reward). */
int spi_b200_fim_contract(spi_b200_model* model, const float* hist, const unsigned char* live, int T, int M, int P,
                          float delta, int accumulate, float* out_JtJ, float* out_trace, void* cuda_stream);

/* CEM elite selection + refit on device.
 *   params [C,P], cost [C] (weighted), n_elite; out_mean [P], out_std [P] updated in place:
 *   mean <- (1-alpha)*mean + alpha*elite_mean, std likewise; out_best [P+1] = best params, cost */
int spi_b200_cem_refit(spi_b200_model* model, const float* params, const float* cost,
                       int C, int P, int n_elite, float alpha, const float* std_floor,
                       float* mean, float* std, float* out_best, void* cuda_stream);

/* CEM sampling on device: params[c,p] = clamp(mean[p] + std[p]*N(0,1), lo[p], hi[p]),
 * counter-based RNG keyed by (seed, iteration, global candidate index c0+c) so that any
 * sharding of the candidates over ranks draws the same population.                            */
int spi_b200_cem_sample(spi_b200_model* model, const float* mean, const float* std,
                        const float* lo, const float* hi, int C, int P, int c0,
                        unsigned long long seed, int iteration, float* out_params,
                        void* cuda_stream);

/* Weighted total cost: total[c] = w0*cost[c,0] + w1*cost[c,1] + w2*cost[c,2]
 * (mass_landscape.py:162-164, mass_opt.py:62-76).                                             */
int spi_b200_weighted_cost(spi_b200_model* model, const float* cost3, int C,
                           float w_pos, float w_quat, float w_joint, float* out_total,
                           void* cuda_stream);

/* FP32 FMA-pipe peak microbenchmark (roofline denominator): returns the
 * measured TFLOP/s of a register-resident FFMA loop on the current device.                    */
int spi_b200_fp32_peak(int iters, float* out_tflops, float* out_ms, void* cuda_stream);

/* Rollout-kernel selection for this handle: 0 = automatic (the warp-specialised Go2-family fast path when the
 * blob has that structure, else the generic leg-per-lane kernel), 1 = force the generic kernel, 2 = require the
 * fast path (error if the blob does not qualify), 3 = the fast path with every candidate's segments padded to whole
 * 32-rollout CTAs (the default gives a candidate its S / 32 full CTAs and packs the left-over segments of several candidates
 * into shared CTAs: S = 1730 -> 54.06 instead of 55 CTAs per candidate; the costs are bit-identical either way — kept for
 * that comparison).  All are CUDA kernels; there is no CPU path. */
enum { SPI_KERNEL_AUTO = 0, SPI_KERNEL_LANE = 1, SPI_KERNEL_WS = 2, SPI_KERNEL_WS_PADDED = 3 };
int spi_b200_model_set_kernel(spi_b200_model* model, int kernel);

/* Number of kernels this library has launched since load (bench.py's `gpu_launches`). */
long long spi_b200_launch_count(void);

/* Roofline instrumentation: when enabled, every launch of the rollout kernel is
 * bracketed by a pair of CUDA events recorded on the launching stream.  spi_b200_timing_read
 * synchronises on the recorded events and returns the summed kernel time and the launch count
 * since the last reset; `reset` != 0 clears the accumulator.  Off by default (zero overhead).   */
int spi_b200_timing_enable(spi_b200_model* model, int enable);
int spi_b200_timing_read(spi_b200_model* model, double* out_total_ms, long long* out_launches, int reset);

#ifdef __cplusplus
}
#endif
#endif /* SPI_B200_H */
