// aba_leg.cuh — leg-per-lane floating-base articulated-body dynamics for the Go2 (sm_100a).
//
// Mapping: one rollout = 4 consecutive lanes of a warp, lane k owns leg k
// (FL, FR, RL, RR) = a 3-joint chain hip(x) - thigh(y) - calf(y); 8 rollouts per warp.  All leg work
// (outward velocity pass, bias forces, foot contact, inward articulated-inertia pass, outward
// acceleration pass) is lane-private, in registers, with compile-time joint axes so every joint
// rotation is a planar rotation.  The four hips' contributions to the base (21 + 6 floats) are summed
// with two xor-butterfly shuffle rounds, after which every lane of the group solves the same 6x6
// system (bitwise identical inputs -> bitwise identical base state on the 4 lanes, no broadcast).
//
// Physics follows the same published algorithm as the CPU oracle (Featherstone, RBDA 2008, Table 9.4)
// but is written independently: symmetric storage, zero rows of the projected inertia skipped,
// chain-composed foot kinematics.
#pragma once
#include <cuda_runtime.h>

#include "../../include/spi_b200.h"

namespace spi {

#define SPI_DEV __device__ __forceinline__

// symmetric 3x3 stored as (xx, yy, zz, xy, xz, yz)
SPI_DEV constexpr int sidx(int i, int j) {
  return (i == j) ? i : ((i + j == 1) ? 3 : ((i + j == 2) ? 4 : 5));
}

struct Spatial6 { float a[3]; float l[3]; };          // motion [w; v] or force [n; f]
struct ABInertia { float I[6]; float H[9]; float M[6]; };  // [[I, H], [H^T, M]], I and M symmetric

// per-lane (per-leg) constants, read once from the device model
struct LegConst {
  float m[3];        // body masses (hip, thigh, calf+foot)
  float h[3][3];     // m * com
  float Io[3][6];    // inertia about the link origin, symmetric storage
  float r[3][3];     // joint origin in the parent frame
  float foot[3];     // foot sphere centre in the calf frame
  float qdef[3], tlim[3];
};

struct SimConst {
  float dt, gz, action_scale, action_clip, kn, cn, mu, dtan, radius, veps2;
  int nsub;
};

// device-resident model (built by spi_b200_model_create from the host blob)
struct DeviceModel {
  SimConst sim;
  float base_inertial[10];
  float lumps[2][10];
  LegConst leg[4];
  float kp[12], kd[12];
};

SPI_DEV void cross3(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
SPI_DEV float cross_comp(const float* a, const float* b, int i) {
  const int j = (i + 1) % 3, k = (i + 2) % 3;
  return a[j] * b[k] - a[k] * b[j];
}
SPI_DEV void sym_mulv(const float* S, const float* v, float* o) {
  o[0] = S[0] * v[0] + S[3] * v[1] + S[4] * v[2];
  o[1] = S[3] * v[0] + S[1] * v[1] + S[5] * v[2];
  o[2] = S[4] * v[0] + S[5] * v[1] + S[2] * v[2];
}

// planar rotation helpers for a joint about coordinate axis AX with (a, b, c) cyclic:
//   R e_a = e_a, R e_b = cs e_b + sn e_c, R e_c = -sn e_b + cs e_c      (child -> parent)
template <int AX> struct Ax {
  static constexpr int a = AX, b = (AX + 1) % 3, c = (AX + 2) % 3;
};
template <int AX> SPI_DEV void rot_up(float cs, float sn, const float* v, float* o) {  // R v
  o[Ax<AX>::a] = v[Ax<AX>::a];
  o[Ax<AX>::b] = cs * v[Ax<AX>::b] - sn * v[Ax<AX>::c];
  o[Ax<AX>::c] = sn * v[Ax<AX>::b] + cs * v[Ax<AX>::c];
}
template <int AX> SPI_DEV void rot_down(float cs, float sn, const float* v, float* o) {  // R^T v
  o[Ax<AX>::a] = v[Ax<AX>::a];
  o[Ax<AX>::b] = cs * v[Ax<AX>::b] + sn * v[Ax<AX>::c];
  o[Ax<AX>::c] = cs * v[Ax<AX>::c] - sn * v[Ax<AX>::b];
}

// quantities a joint keeps between the passes
struct JointKeep {
  float cs, sn;
  float cab, cac, clb, clc;  // velocity-product acceleration c = v x (e_a qd): components b, c
  float Ua[3], Ul[3], dinv, u;
};

// ---- outward pass for one joint: velocity, c, bias force of the rigid body -----------------------
template <int AX>
SPI_DEV void joint_outward(const Spatial6& vp, const float* r, float q, float qd, float mass, const float* h,
                           const float* Io, Spatial6& v, JointKeep& k, Spatial6& pA) {
  constexpr int a = Ax<AX>::a, b = Ax<AX>::b, c = Ax<AX>::c;
  sincosf(q, &k.sn, &k.cs);
  float t[3], wxr[3];
  rot_down<AX>(k.cs, k.sn, vp.a, v.a);
  v.a[a] += qd;
  cross3(vp.a, r, wxr);
  t[0] = vp.l[0] + wxr[0]; t[1] = vp.l[1] + wxr[1]; t[2] = vp.l[2] + wxr[2];
  rot_down<AX>(k.cs, k.sn, t, v.l);
  k.cab = v.a[c] * qd;  k.cac = -(v.a[b] * qd);
  k.clb = v.l[c] * qd;  k.clc = -(v.l[b] * qd);
  // momentum: n = Io w + h x v,  f = m v - h x w ;  pA = [w x n + v x f ; w x f]
  float n[3], f[3], hv[3], hw[3], t1[3], t2[3];
  sym_mulv(Io, v.a, n);
  cross3(h, v.l, hv);
  cross3(h, v.a, hw);
#pragma unroll
  for (int i = 0; i < 3; i++) { n[i] += hv[i]; f[i] = mass * v.l[i] - hw[i]; }
  cross3(v.a, n, t1);
  cross3(v.l, f, t2);
  cross3(v.a, f, pA.l);
#pragma unroll
  for (int i = 0; i < 3; i++) pA.a[i] = t1[i] + t2[i];
}

// rigid-body inertia -> articulated-inertia storage
SPI_DEV void abi_from_rigid(float mass, const float* h, const float* Io, ABInertia& A) {
#pragma unroll
  for (int i = 0; i < 6; i++) A.I[i] = Io[i];
  // H = skew(h)
  A.H[0] = 0.f;   A.H[1] = -h[2]; A.H[2] = h[1];
  A.H[3] = h[2];  A.H[4] = 0.f;   A.H[5] = -h[0];
  A.H[6] = -h[1]; A.H[7] = h[0];  A.H[8] = 0.f;
  A.M[0] = A.M[1] = A.M[2] = mass;
  A.M[3] = A.M[4] = A.M[5] = 0.f;
}

// ---- inward pass for one joint: project out the joint, transform to the parent, accumulate ---------
// IAp / pAp must already hold the parent's own inertia / bias force.
template <int AX>
SPI_DEV void joint_inward(const ABInertia& IA, const Spatial6& pA, float tau, const float* r, JointKeep& k,
                          ABInertia& IAp, Spatial6& pAp) {
  constexpr int a = Ax<AX>::a, b = Ax<AX>::b, c = Ax<AX>::c;
  const float cs = k.cs, sn = k.sn;
#pragma unroll
  for (int i = 0; i < 3; i++) { k.Ua[i] = IA.I[sidx(i, a)]; k.Ul[i] = IA.H[3 * a + i]; }
  k.dinv = 1.0f / k.Ua[a];
  k.u = tau - pA.a[a];
  float Uad[3], Uld[3];
#pragma unroll
  for (int i = 0; i < 3; i++) { Uad[i] = k.Ua[i] * k.dinv; Uld[i] = k.Ul[i] * k.dinv; }
  // projected inertia Ia = IA - U U^T / D.  Row/column a of I and row a of H vanish identically.
  const float Ibb = IA.I[sidx(b, b)] - Uad[b] * k.Ua[b];
  const float Ibc = IA.I[sidx(b, c)] - Uad[b] * k.Ua[c];
  const float Icc = IA.I[sidx(c, c)] - Uad[c] * k.Ua[c];
  float Hb[3], Hc[3], Ma[6];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    Hb[j] = IA.H[3 * b + j] - Uad[b] * k.Ul[j];
    Hc[j] = IA.H[3 * c + j] - Uad[c] * k.Ul[j];
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = i; j < 3; j++) Ma[sidx(i, j)] = IA.M[sidx(i, j)] - Uld[i] * k.Ul[j];
  // pa = pA + Ia c + U u / D     (c has no component along a)
  const float ud = k.u * k.dinv;
  float pa_a[3], pa_l[3];
  pa_a[a] = pA.a[a] + ud * k.Ua[a];
  pa_a[b] = pA.a[b] + Ibb * k.cab + Ibc * k.cac + Hb[b] * k.clb + Hb[c] * k.clc + ud * k.Ua[b];
  pa_a[c] = pA.a[c] + Ibc * k.cab + Icc * k.cac + Hc[b] * k.clb + Hc[c] * k.clc + ud * k.Ua[c];
#pragma unroll
  for (int j = 0; j < 3; j++)
    pa_l[j] = pA.l[j] + Hb[j] * k.cab + Hc[j] * k.cac + Ma[sidx(j, b)] * k.clb + Ma[sidx(j, c)] * k.clc + ud * k.Ul[j];
  // rotate the blocks to parent orientation:  X' = R X R^T
  //   I: only the (b,c) 2x2 block is non-zero
  const float t1 = cs * Ibb - sn * Ibc, t2 = cs * Ibc - sn * Icc;
  const float t3 = sn * Ibb + cs * Ibc, t4 = sn * Ibc + cs * Icc;
  const float I2bb = t1 * cs - t2 * sn, I2bc = t1 * sn + t2 * cs, I2cc = t3 * sn + t4 * cs;
  //   H: rows b, c non-zero
  float Xb[3], Xc[3], H2[9];
#pragma unroll
  for (int j = 0; j < 3; j++) { Xb[j] = cs * Hb[j] - sn * Hc[j]; Xc[j] = sn * Hb[j] + cs * Hc[j]; }
  H2[3 * a + 0] = H2[3 * a + 1] = H2[3 * a + 2] = 0.f;
  H2[3 * b + a] = Xb[a]; H2[3 * b + b] = cs * Xb[b] - sn * Xb[c]; H2[3 * b + c] = sn * Xb[b] + cs * Xb[c];
  H2[3 * c + a] = Xc[a]; H2[3 * c + b] = cs * Xc[b] - sn * Xc[c]; H2[3 * c + c] = sn * Xc[b] + cs * Xc[c];
  //   M: full symmetric
  float M2[6];
  {
    const float Bbb = cs * Ma[sidx(b, b)] - sn * Ma[sidx(c, b)], Bbc = cs * Ma[sidx(b, c)] - sn * Ma[sidx(c, c)];
    const float Bcb = sn * Ma[sidx(b, b)] + cs * Ma[sidx(c, b)], Bcc = sn * Ma[sidx(b, c)] + cs * Ma[sidx(c, c)];
    M2[sidx(a, a)] = Ma[sidx(a, a)];
    M2[sidx(a, b)] = cs * Ma[sidx(a, b)] - sn * Ma[sidx(a, c)];
    M2[sidx(a, c)] = sn * Ma[sidx(a, b)] + cs * Ma[sidx(a, c)];
    M2[sidx(b, b)] = cs * Bbb - sn * Bbc;
    M2[sidx(b, c)] = sn * Bbb + cs * Bbc;
    M2[sidx(c, c)] = sn * Bcb + cs * Bcc;
  }
  // shift the reference point by r:  Hp = H2 + r x M2 ;  Ip = I2 + r x H2^T - Hp r x
  float Hp[9];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
    for (int j = 0; j < 3; j++) Hp[3 * i + j] = H2[3 * i + j] + r[i1] * M2[sidx(i2, j)] - r[i2] * M2[sidx(i1, j)];
  }
  float I2[6];
  I2[sidx(a, a)] = 0.f; I2[sidx(a, b)] = 0.f; I2[sidx(a, c)] = 0.f;
  I2[sidx(b, b)] = I2bb; I2[sidx(b, c)] = I2bc; I2[sidx(c, c)] = I2cc;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = i; j < 3; j++) {
      const float v = I2[sidx(i, j)] + cross_comp(r, &H2[3 * j], i) - cross_comp(&Hp[3 * i], r, j);
      IAp.I[sidx(i, j)] += v;
    }
#pragma unroll
  for (int i = 0; i < 9; i++) IAp.H[i] += Hp[i];
#pragma unroll
  for (int i = 0; i < 6; i++) IAp.M[i] += M2[i];
  // force to the parent
  float fl[3], fa[3], rxf[3];
  rot_up<AX>(cs, sn, pa_l, fl);
  rot_up<AX>(cs, sn, pa_a, fa);
  cross3(r, fl, rxf);
#pragma unroll
  for (int i = 0; i < 3; i++) { pAp.a[i] += fa[i] + rxf[i]; pAp.l[i] += fl[i]; }
}

// ---- outward acceleration pass for one joint -------------------------------------------------------
template <int AX>
SPI_DEV float joint_accel(const Spatial6& ap, const float* r, const JointKeep& k, Spatial6& acc) {
  constexpr int a = Ax<AX>::a, b = Ax<AX>::b, c = Ax<AX>::c;
  float t[3], axr[3];
  rot_down<AX>(k.cs, k.sn, ap.a, acc.a);
  cross3(ap.a, r, axr);
  t[0] = ap.l[0] + axr[0]; t[1] = ap.l[1] + axr[1]; t[2] = ap.l[2] + axr[2];
  rot_down<AX>(k.cs, k.sn, t, acc.l);
  acc.a[b] += k.cab; acc.a[c] += k.cac;
  acc.l[b] += k.clb; acc.l[c] += k.clc;
  const float dotU = k.Ua[0] * acc.a[0] + k.Ua[1] * acc.a[1] + k.Ua[2] * acc.a[2] +
                     k.Ul[0] * acc.l[0] + k.Ul[1] * acc.l[1] + k.Ul[2] * acc.l[2];
  const float qdd = (k.u - dotU) * k.dinv;
  acc.a[a] += qdd;
  return qdd;
}

// 6x6 SPD solve (LDL^T), A given as the articulated inertia blocks, b = -pA
SPI_DEV void solve_base(const ABInertia& A, const Spatial6& pA, Spatial6& a0) {
  float M[6][6];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      M[i][j] = A.I[sidx(i, j)];
      M[i][3 + j] = A.H[3 * i + j];
      M[3 + i][j] = A.H[3 * j + i];
      M[3 + i][3 + j] = A.M[sidx(i, j)];
    }
  float L[6][6], D[6], Dinv[6];
#pragma unroll
  for (int j = 0; j < 6; j++) {
    float d = M[j][j];
#pragma unroll
    for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k] * D[k];
    D[j] = d;
    Dinv[j] = 1.0f / d;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      float s = M[i][j];
#pragma unroll
      for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k] * D[k];
      L[i][j] = s * Dinv[j];
    }
  }
  float y[6], x[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float s = (i < 3) ? -pA.a[i] : -pA.l[i - 3];
#pragma unroll
    for (int k = 0; k < i; k++) s -= L[i][k] * y[k];
    y[i] = s;
  }
#pragma unroll
  for (int i = 0; i < 6; i++) y[i] *= Dinv[i];
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    float s = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) s -= L[k][i] * x[k];
    x[i] = s;
  }
#pragma unroll
  for (int i = 0; i < 3; i++) { a0.a[i] = x[i]; a0.l[i] = x[3 + i]; }
}

// state of one rollout as seen by one lane: the base (replicated on the 4 lanes) + its leg's joints
struct LaneState {
  float p[3], quat[4], v[3], w[3];
  float q[3], qd[3];
};

// rigid inertia of the base for this candidate: mass, h = m c, Io (about the base origin)
struct BaseInertia { float m; float h[3]; float Io[6]; };

SPI_DEV void add_point_inertia(BaseInertia& B, float mass, const float* c, const float* Ic) {
  const float cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
  B.m += mass;
#pragma unroll
  for (int i = 0; i < 3; i++) B.h[i] += mass * c[i];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = i; j < 3; j++)
      B.Io[sidx(i, j)] += Ic[sidx(i, j)] + mass * ((i == j ? cc : 0.f) - c[i] * c[j]);
}

// sum over the 4 lanes of a rollout group (xor butterfly: every lane ends with the same bits)
SPI_DEV float group_sum(float x) {
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  x += __shfl_xor_sync(0xffffffffu, x, 2);
  return x;
}

// ---- one integrator sub-step of length h under joint torques tau[3] (this lane's leg) --------------
// foot_force (optional): world-frame contact force on this lane's foot.
// ext (optional): external [torque; force] on this lane's three leg bodies (ext[0..18)) and on the base (ext_base[6]), each in
// the body's link frame about its origin — spi_b200_sim_step_ext.
SPI_DEV void substep(const SimConst& S, const LegConst& L, const BaseInertia& B, LaneState& s, const float* tau,
                     float h, float* foot_force, const float* ext = nullptr, const float* ext_base = nullptr) {
  // base rotation (body -> world) and body-frame velocities
  float R[9];
  {
    const float x = s.quat[0], y = s.quat[1], z = s.quat[2], w = s.quat[3];
    const float xx = 2.f * x * x, yy = 2.f * y * y, zz = 2.f * z * z;
    const float xy = 2.f * x * y, xz = 2.f * x * z, yz = 2.f * y * z;
    const float wx = 2.f * w * x, wy = 2.f * w * y, wz = 2.f * w * z;
    R[0] = 1.f - (yy + zz); R[1] = xy - wz;         R[2] = xz + wy;
    R[3] = xy + wz;         R[4] = 1.f - (xx + zz); R[5] = yz - wx;
    R[6] = xz - wy;         R[7] = yz + wx;         R[8] = 1.f - (xx + yy);
  }
  Spatial6 v0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    v0.a[i] = R[i] * s.w[0] + R[3 + i] * s.w[1] + R[6 + i] * s.w[2];
    v0.l[i] = R[i] * s.v[0] + R[3 + i] * s.v[1] + R[6 + i] * s.v[2];
  }
  // outward pass down the leg
  Spatial6 v1, v2, v3, p1, p2, p3;
  JointKeep k1, k2, k3;
  joint_outward<0>(v0, L.r[0], s.q[0], s.qd[0], L.m[0], L.h[0], L.Io[0], v1, k1, p1);
  joint_outward<1>(v1, L.r[1], s.q[1], s.qd[1], L.m[1], L.h[1], L.Io[1], v2, k2, p2);
  joint_outward<1>(v2, L.r[2], s.q[2], s.qd[2], L.m[2], L.h[2], L.Io[2], v3, k3, p3);
  // foot contact (compliant sphere on the plane z = 0)
  {
    // foot centre in base coordinates through the chain
    float t3[3], t2[3], t1[3], u3[3], u2[3];
    rot_up<1>(k3.cs, k3.sn, L.foot, t3);
#pragma unroll
    for (int i = 0; i < 3; i++) u3[i] = L.r[2][i] + t3[i];
    rot_up<1>(k2.cs, k2.sn, u3, t2);
#pragma unroll
    for (int i = 0; i < 3; i++) u2[i] = L.r[1][i] + t2[i];
    rot_up<0>(k1.cs, k1.sn, u2, t1);
    const float fbx = L.r[0][0] + t1[0], fby = L.r[0][1] + t1[1], fbz = L.r[0][2] + t1[2];
    const float pz = s.p[2] + R[6] * fbx + R[7] * fby + R[8] * fbz;
    const float depth = S.radius - pz;
    float F[3] = {0.f, 0.f, 0.f};
    if (depth > 0.f) {
      // foot-centre velocity: calf frame -> base -> world
      float wxo[3], vc[3], a2[3], a1[3], vb[3], vw[3];
      cross3(v3.a, L.foot, wxo);
#pragma unroll
      for (int i = 0; i < 3; i++) vc[i] = v3.l[i] + wxo[i];
      rot_up<1>(k3.cs, k3.sn, vc, a2);
      rot_up<1>(k2.cs, k2.sn, a2, a1);
      rot_up<0>(k1.cs, k1.sn, a1, vb);
#pragma unroll
      for (int i = 0; i < 3; i++) vw[i] = R[3 * i] * vb[0] + R[3 * i + 1] * vb[1] + R[3 * i + 2] * vb[2];
      float fn = S.kn * depth * (1.f - S.cn * vw[2]);
      fn = fmaxf(fn, 0.f);
      const float speed = sqrtf(vw[0] * vw[0] + vw[1] * vw[1] + S.veps2);
      const float coef = fminf(S.dtan, S.mu * fn / speed);
      F[0] = -(coef * vw[0]); F[1] = -(coef * vw[1]); F[2] = fn;
      // world -> base -> calf
      float fb[3], g1[3], g2[3], fc[3], nc[3];
#pragma unroll
      for (int i = 0; i < 3; i++) fb[i] = R[i] * F[0] + R[3 + i] * F[1] + R[6 + i] * F[2];
      rot_down<0>(k1.cs, k1.sn, fb, g1);
      rot_down<1>(k2.cs, k2.sn, g1, g2);
      rot_down<1>(k3.cs, k3.sn, g2, fc);
      cross3(L.foot, fc, nc);
#pragma unroll
      for (int i = 0; i < 3; i++) { p3.a[i] -= nc[i]; p3.l[i] -= fc[i]; }
    }
    if (foot_force) { foot_force[0] = F[0]; foot_force[1] = F[1]; foot_force[2] = F[2]; }
  }
  if (ext) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
      p1.a[i] -= ext[i]; p1.l[i] -= ext[3 + i];
      p2.a[i] -= ext[6 + i]; p2.l[i] -= ext[9 + i];
      p3.a[i] -= ext[12 + i]; p3.l[i] -= ext[15 + i];
    }
  }
  // inward pass up the leg
  ABInertia A3, A2, A1, A0;
  abi_from_rigid(L.m[2], L.h[2], L.Io[2], A3);
  abi_from_rigid(L.m[1], L.h[1], L.Io[1], A2);
  joint_inward<1>(A3, p3, tau[2], L.r[2], k3, A2, p2);
  abi_from_rigid(L.m[0], L.h[0], L.Io[0], A1);
  joint_inward<1>(A2, p2, tau[1], L.r[1], k2, A1, p1);
  // hip -> base contribution of this leg, then sum over the 4 legs
  Spatial6 p0;
#pragma unroll
  for (int i = 0; i < 6; i++) { A0.I[i] = 0.f; A0.M[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < 9; i++) A0.H[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 3; i++) { p0.a[i] = 0.f; p0.l[i] = 0.f; }
  joint_inward<0>(A1, p1, tau[0], L.r[0], k1, A0, p0);
#pragma unroll
  for (int i = 0; i < 6; i++) { A0.I[i] = group_sum(A0.I[i]); A0.M[i] = group_sum(A0.M[i]); }
#pragma unroll
  for (int i = 0; i < 9; i++) A0.H[i] = group_sum(A0.H[i]);
#pragma unroll
  for (int i = 0; i < 3; i++) { p0.a[i] = group_sum(p0.a[i]); p0.l[i] = group_sum(p0.l[i]); }
  // base's own inertia and bias force
  {
#pragma unroll
    for (int i = 0; i < 6; i++) A0.I[i] += B.Io[i];
    A0.H[1] -= B.h[2]; A0.H[2] += B.h[1];
    A0.H[3] += B.h[2]; A0.H[5] -= B.h[0];
    A0.H[6] -= B.h[1]; A0.H[7] += B.h[0];
    A0.M[0] += B.m; A0.M[1] += B.m; A0.M[2] += B.m;
    float n[3], f[3], hv[3], hw[3], t1[3], t2[3], t3[3];
    sym_mulv(B.Io, v0.a, n);
    cross3(B.h, v0.l, hv);
    cross3(B.h, v0.a, hw);
#pragma unroll
    for (int i = 0; i < 3; i++) { n[i] += hv[i]; f[i] = B.m * v0.l[i] - hw[i]; }
    cross3(v0.a, n, t1);
    cross3(v0.l, f, t2);
    cross3(v0.a, f, t3);
#pragma unroll
    for (int i = 0; i < 3; i++) { p0.a[i] += t1[i] + t2[i]; p0.l[i] += t3[i]; }
  }
  if (ext_base) {
#pragma unroll
    for (int i = 0; i < 3; i++) { p0.a[i] -= ext_base[i]; p0.l[i] -= ext_base[3 + i]; }
  }
  Spatial6 a0;
  solve_base(A0, p0, a0);
  // outward acceleration pass
  Spatial6 a1, a2, a3;
  const float qdd0 = joint_accel<0>(a0, L.r[0], k1, a1);
  const float qdd1 = joint_accel<1>(a1, L.r[1], k2, a2);
  const float qdd2 = joint_accel<1>(a2, L.r[2], k3, a3);
  // semi-implicit Euler
  s.qd[0] += h * qdd0; s.q[0] += h * s.qd[0];
  s.qd[1] += h * qdd1; s.q[1] += h * s.qd[1];
  s.qd[2] += h * qdd2; s.q[2] += h * s.qd[2];
  {
    // gravity enters as a uniform acceleration of every body (RBDA 9.4): a0.l += R^T g
    float accb[3], wxv[3];
    cross3(v0.a, v0.l, wxv);
#pragma unroll
    for (int i = 0; i < 3; i++) accb[i] = a0.l[i] + R[6 + i] * S.gz + wxv[i];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      s.w[i] += h * (R[3 * i] * a0.a[0] + R[3 * i + 1] * a0.a[1] + R[3 * i + 2] * a0.a[2]);
      s.v[i] += h * (R[3 * i] * accb[0] + R[3 * i + 1] * accb[1] + R[3 * i + 2] * accb[2]);
      s.p[i] += h * s.v[i];
    }
    const float hx = 0.5f * h;
    const float x = s.quat[0], y = s.quat[1], z = s.quat[2], w = s.quat[3];
    const float nx = x + hx * (s.w[0] * w + s.w[1] * z - s.w[2] * y);
    const float ny = y + hx * (s.w[1] * w + s.w[2] * x - s.w[0] * z);
    const float nz = z + hx * (s.w[2] * w + s.w[0] * y - s.w[1] * x);
    const float nw = w - hx * (s.w[0] * x + s.w[1] * y + s.w[2] * z);
    const float inv = rsqrtf(nx * nx + ny * ny + nz * nz + nw * nw);
    s.quat[0] = nx * inv; s.quat[1] = ny * inv; s.quat[2] = nz * inv; s.quat[3] = nw * inv;
  }
}

// PD law + torque clip + motor model for this lane's 3 joints
// (legged_robot_base.py:545,557; go2_omni.py:436-437; active_sysid_openloop.py:184-186,356-400)
SPI_DEV void lane_torques(const SimConst& S, const LegConst& L, const float* act /*clipped*/, const float* q,
                          const float* qd, const float* kp, const float* kd, const float* motor, int motor_model,
                          unsigned flags, float* tau) {
#pragma unroll
  for (int j = 0; j < 3; j++) {
    float as = act[j] * S.action_scale;
    if (j == 0 && (flags & SPI_FLAG_HIP_HALF)) as *= 0.5f;
    float t = kp[j] * (as + L.qdef[j] - q[j]) - kd[j] * qd[j];
    const float g = motor[j];
    if (motor_model == SPI_MOTOR_VEC3_TANH && (flags & SPI_FLAG_TANH_BEFORE_CLIP)) {
      t = g * tanhf((1.0f / g) * t);
      t = fminf(fmaxf(t, -L.tlim[j]), L.tlim[j]);
    } else {
      t = fminf(fmaxf(t, -L.tlim[j]), L.tlim[j]);
      if (motor_model == SPI_MOTOR_SCALAR) t *= motor[0];
      else if (motor_model == SPI_MOTOR_VEC3) t *= g;
      else if (motor_model == SPI_MOTOR_VEC3_TANH) t = g * tanhf((1.0f / g) * t);
    }
    tau[j] = t;
  }
}

// candidate row -> base inertia (+ head lumps) and this lane's motor parameters
// (isaacgym_active_sysid.py:61-94 setters; mass_opt.py:158-160 mass_scale; DESIGN.md D8/D15 flags)
struct ParamIds { int n; int id[16]; };

SPI_DEV void apply_candidate(const DeviceModel& M, const float* row, const ParamIds& ids, unsigned flags,
                             BaseInertia& B, float* motor3) {
  float rec[10];
#pragma unroll
  for (int k = 0; k < 10; k++) rec[k] = M.base_inertial[k];
  motor3[0] = motor3[1] = motor3[2] = 20.0f;
  float mass = rec[0];
  if (row) {
    for (int p = 0; p < ids.n; p++) {
      if (ids.id[p] == SPI_PARAM_MASS) mass = row[p];
      if (ids.id[p] == SPI_PARAM_MASS_SCALE) mass = rec[0] * row[p];
    }
  }
  if (!(flags & SPI_FLAG_INERTIA_KEEP)) {
    const float sc = mass / rec[0];
#pragma unroll
    for (int k = 4; k < 10; k++) rec[k] *= sc;
  }
  rec[0] = mass;
  if (row) {
    for (int p = 0; p < ids.n; p++) {
      const float v = row[p];
      const int id = ids.id[p];
      if (id >= SPI_PARAM_COMX && id <= SPI_PARAM_INERTIAYZ) {
        if (id == SPI_PARAM_INERTIAY && (flags & SPI_FLAG_STRICT_INERTIAY)) continue;
        // rec index == param id for ids 1..9 (com xyz, I xx yy zz xy xz yz)
#pragma unroll
        for (int k = 1; k < 10; k++) if (id == k) rec[k] = v;
      } else if (id == SPI_PARAM_MOTOR_HIP) motor3[0] = v;
      else if (id == SPI_PARAM_MOTOR_THIGH) motor3[1] = v;
      else if (id == SPI_PARAM_MOTOR_CALF) motor3[2] = v;
    }
  }
  B.m = 0.f;
#pragma unroll
  for (int i = 0; i < 3; i++) B.h[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 6; i++) B.Io[i] = 0.f;
  add_point_inertia(B, rec[0], rec + 1, rec + 4);
  add_point_inertia(B, M.lumps[0][0], M.lumps[0] + 1, M.lumps[0] + 4);
  add_point_inertia(B, M.lumps[1][0], M.lumps[1] + 1, M.lumps[1] + 4);
}

}  // namespace spi
