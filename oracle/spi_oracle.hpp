// oracle/spi_oracle.hpp
//
// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / `--impl reference` legs may build, link or call anything under oracle/.
//
// CPU restatement of the SPI-Active SysID hot path, templated on the scalar so the
// same source runs in double (the parity oracle), in float (the CPU baseline and the fp32 noise
// floor) and in an op-counting scalar (the algorithmic-FLOP figure of the roofline).
//
// PARITY STATUS
//   * control / replay / cost semantics restate in-tree reference code and are pinned by the
//     golden vectors under tests/golden/ (generated from numpy restatements of the reference torch
//     arithmetic, see tests/golden/make_golden.py):
//       - torque law          spigym/envs/legged_base_task/legged_robot_base.py:185-209, 529-560
//       - hip x0.5            spigym/envs/locomotion/go2_omni.py:436-437
//       - motor models        spigym/envs/sysid/active_sysid_openloop.py:174-187, 356-400
//       - windowing / cost    scripts/eval.py:101-171, 217-310
//       - parameter setters   spigym/simulator/isaacgym/isaacgym_active_sysid.py:61-94
//       - FIM reward          spigym/envs/sysid/active_sysid_openloop.py:402-426
//   * rigid-body physics: "PARITY UNPINNED".  The reference delegates physics to the closed-source
//     Isaac Gym Preview 4 / PhysX binary (not vendored, not pinned in pyproject.toml / uv.lock, not
//     installable here: SURVEY.md §8c) and ships no golden vectors.  This file restates the published
//     algorithm the engine uses instead — Featherstone's floating-base articulated-body algorithm
//     (Rigid Body Dynamics Algorithms, 2008, Table 9.4) in link coordinates, a compliant
//     Hunt-Crossley sphere-plane foot contact with Coulomb-capped viscous friction, semi-implicit
//     Euler — and is validated by physics invariants and an independent dense CRBA/RNEA
//     cross-check in tests/test_oracle_physics.py.
//
// The layout of the model blob, the parameter ids, motor-model ids and flags are those of
// include/spi_b200.h.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

#include "../include/spi_b200.h"

namespace spi_oracle {

// ----------------------------------------------------------------------------------------------
// op-counting scalar
// ----------------------------------------------------------------------------------------------
struct OpCounts {
  uint64_t add = 0, mul = 0, div = 0, sqrt = 0, trans = 0, cmp = 0;
  uint64_t flops() const { return add + mul + div + sqrt + trans + cmp; }
};
inline OpCounts& op_counts() {
  static thread_local OpCounts c;
  return c;
}
struct Counted {
  double v;
  Counted() : v(0) {}
  Counted(double x) : v(x) {}
  explicit operator double() const { return v; }
};
inline Counted operator+(Counted a, Counted b) { op_counts().add++; return Counted(a.v + b.v); }
inline Counted operator-(Counted a, Counted b) { op_counts().add++; return Counted(a.v - b.v); }
inline Counted operator-(Counted a) { return Counted(-a.v); }
inline Counted operator*(Counted a, Counted b) { op_counts().mul++; return Counted(a.v * b.v); }
inline Counted operator/(Counted a, Counted b) { op_counts().div++; return Counted(a.v / b.v); }
inline Counted& operator+=(Counted& a, Counted b) { a = a + b; return a; }
inline Counted& operator-=(Counted& a, Counted b) { a = a - b; return a; }
inline Counted& operator*=(Counted& a, Counted b) { a = a * b; return a; }
inline bool operator>(Counted a, Counted b) { op_counts().cmp++; return a.v > b.v; }
inline bool operator<(Counted a, Counted b) { op_counts().cmp++; return a.v < b.v; }

template <class T> inline T s_sqrt(T x) { return std::sqrt(x); }
template <> inline Counted s_sqrt(Counted x) { op_counts().sqrt++; return Counted(std::sqrt(x.v)); }
template <class T> inline T s_tanh(T x) { return std::tanh(x); }
template <> inline Counted s_tanh(Counted x) { op_counts().trans++; return Counted(std::tanh(x.v)); }
template <class T> inline T s_sin(T x) { return std::sin(x); }
template <> inline Counted s_sin(Counted x) { op_counts().trans++; return Counted(std::sin(x.v)); }
template <class T> inline T s_cos(T x) { return std::cos(x); }
template <> inline Counted s_cos(Counted x) { op_counts().trans++; return Counted(std::cos(x.v)); }
template <class T> inline T s_min(T a, T b) { return (a < b) ? a : b; }
template <class T> inline T s_max(T a, T b) { return (a > b) ? a : b; }
template <class T> inline T s_clip(T x, T lo, T hi) { return s_min(s_max(x, lo), hi); }
template <class T> inline double to_double(T x) { return (double)x; }
template <class T> inline bool s_finite(T x) { return std::isfinite(to_double(x)); }

// ----------------------------------------------------------------------------------------------
// 3-vectors / 3x3 matrices
// ----------------------------------------------------------------------------------------------
template <class T> struct V3 {
  T x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(T a, T b, T c) : x(a), y(b), z(c) {}
  T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <class T> inline V3<T> operator+(const V3<T>& a, const V3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> inline V3<T> operator-(const V3<T>& a, const V3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> inline V3<T> operator*(T s, const V3<T>& a) { return {s * a.x, s * a.y, s * a.z}; }
template <class T> inline T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> inline V3<T> cross(const V3<T>& a, const V3<T>& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

template <class T> struct M3 {
  T m[3][3];
  M3() { for (auto& r : m) for (auto& e : r) e = T(0); }
};
template <class T> inline V3<T> mul(const M3<T>& A, const V3<T>& v) {
  return {A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z,
          A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
          A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z};
}
template <class T> inline V3<T> mulT(const M3<T>& A, const V3<T>& v) {  // A^T v
  return {A.m[0][0] * v.x + A.m[1][0] * v.y + A.m[2][0] * v.z,
          A.m[0][1] * v.x + A.m[1][1] * v.y + A.m[2][1] * v.z,
          A.m[0][2] * v.x + A.m[1][2] * v.y + A.m[2][2] * v.z};
}
template <class T> inline M3<T> mul(const M3<T>& A, const M3<T>& B) {
  M3<T> C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
  return C;
}
template <class T> inline M3<T> transpose(const M3<T>& A) {
  M3<T> C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C.m[i][j] = A.m[j][i];
  return C;
}
template <class T> inline M3<T> add(const M3<T>& A, const M3<T>& B) {
  M3<T> C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C.m[i][j] = A.m[i][j] + B.m[i][j];
  return C;
}
template <class T> inline M3<T> sub(const M3<T>& A, const M3<T>& B) {
  M3<T> C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C.m[i][j] = A.m[i][j] - B.m[i][j];
  return C;
}
template <class T> inline M3<T> skew(const V3<T>& v) {
  M3<T> S;
  S.m[0][1] = -v.z; S.m[0][2] = v.y;
  S.m[1][0] = v.z;  S.m[1][2] = -v.x;
  S.m[2][0] = -v.y; S.m[2][1] = v.x;
  return S;
}
// r x A  (each column crossed) and A rx
template <class T> inline M3<T> cross_left(const V3<T>& r, const M3<T>& A) {
  M3<T> C;
  for (int j = 0; j < 3; j++) {
    V3<T> col(A.m[0][j], A.m[1][j], A.m[2][j]);
    V3<T> c = cross(r, col);
    C.m[0][j] = c.x; C.m[1][j] = c.y; C.m[2][j] = c.z;
  }
  return C;
}
template <class T> inline M3<T> cross_right(const M3<T>& A, const V3<T>& r) {  // A * skew(r)
  // row_i(A) x-> (a × r) with sign: (A rx)_i = -(r × a_i)^T = (a_i × r)^T
  M3<T> C;
  for (int i = 0; i < 3; i++) {
    V3<T> row(A.m[i][0], A.m[i][1], A.m[i][2]);
    V3<T> c = cross(row, r);
    C.m[i][0] = c.x; C.m[i][1] = c.y; C.m[i][2] = c.z;
  }
  return C;
}

// rotation of angle q about coordinate axis `ax` (child -> parent coordinates)
template <class T> inline M3<T> axis_rotation(int ax, T q) {
  T c = s_cos(q), s = s_sin(q);
  M3<T> R;
  R.m[0][0] = R.m[1][1] = R.m[2][2] = T(1);
  if (ax == 0) { R.m[1][1] = c; R.m[1][2] = -s; R.m[2][1] = s; R.m[2][2] = c; }
  else if (ax == 1) { R.m[0][0] = c; R.m[0][2] = s; R.m[2][0] = -s; R.m[2][2] = c; }
  else { R.m[0][0] = c; R.m[0][1] = -s; R.m[1][0] = s; R.m[1][1] = c; }
  return R;
}
// unit quaternion (x,y,z,w) -> rotation matrix (body -> world)
template <class T> inline M3<T> quat_to_matrix(const T* q) {
  T x = q[0], y = q[1], z = q[2], w = q[3];
  T two(2);
  T xx = two * x * x, yy = two * y * y, zz = two * z * z;
  T xy = two * x * y, xz = two * x * z, yz = two * y * z;
  T wx = two * w * x, wy = two * w * y, wz = two * w * z;
  M3<T> R;
  R.m[0][0] = T(1) - (yy + zz); R.m[0][1] = xy - wz;          R.m[0][2] = xz + wy;
  R.m[1][0] = xy + wz;          R.m[1][1] = T(1) - (xx + zz); R.m[1][2] = yz - wx;
  R.m[2][0] = xz - wy;          R.m[2][1] = yz + wx;          R.m[2][2] = T(1) - (xx + yy);
  return R;
}


// ---- rotations about a coordinate axis, exploiting that only the (b,c) plane mixes ----------------
// R = rot(axis a, angle): R e_a = e_a, R e_b = c e_b + s e_c, R e_c = -s e_b + c e_c, (a,b,c) cyclic.
template <class T> struct AxRot { int a, b, c; T cs, sn; };
template <class T> inline AxRot<T> make_axrot(int ax, T q) {
  AxRot<T> r; r.a = ax; r.b = (ax + 1) % 3; r.c = (ax + 2) % 3; r.cs = s_cos(q); r.sn = s_sin(q); return r;
}
template <class T> inline V3<T> rot(const AxRot<T>& r, const V3<T>& v) {   // R v   (child -> parent)
  V3<T> o; o[r.a] = v[r.a];
  o[r.b] = r.cs * v[r.b] - r.sn * v[r.c];
  o[r.c] = r.sn * v[r.b] + r.cs * v[r.c];
  return o;
}
template <class T> inline V3<T> rotT(const AxRot<T>& r, const V3<T>& v) {  // R^T v (parent -> child)
  V3<T> o; o[r.a] = v[r.a];
  o[r.b] = r.cs * v[r.b] + r.sn * v[r.c];
  o[r.c] = r.cs * v[r.c] - r.sn * v[r.b];
  return o;
}
// A R  (compose body->world rotation with the joint rotation): only columns b, c change
template <class T> inline M3<T> mul_rot(const M3<T>& A, const AxRot<T>& r) {
  M3<T> o;
  for (int i = 0; i < 3; i++) {
    o.m[i][r.a] = A.m[i][r.a];
    o.m[i][r.b] = r.cs * A.m[i][r.b] + r.sn * A.m[i][r.c];
    o.m[i][r.c] = r.cs * A.m[i][r.c] - r.sn * A.m[i][r.b];
  }
  return o;
}
// R A R^T for a general 3x3
template <class T> inline M3<T> rot_gen(const AxRot<T>& r, const M3<T>& A) {
  M3<T> B, o;
  for (int j = 0; j < 3; j++) {
    B.m[r.a][j] = A.m[r.a][j];
    B.m[r.b][j] = r.cs * A.m[r.b][j] - r.sn * A.m[r.c][j];
    B.m[r.c][j] = r.sn * A.m[r.b][j] + r.cs * A.m[r.c][j];
  }
  for (int i = 0; i < 3; i++) {
    o.m[i][r.a] = B.m[i][r.a];
    o.m[i][r.b] = r.cs * B.m[i][r.b] - r.sn * B.m[i][r.c];
    o.m[i][r.c] = r.sn * B.m[i][r.b] + r.cs * B.m[i][r.c];
  }
  return o;
}
// R A R^T for a symmetric 3x3 (upper triangle computed, mirrored)
template <class T> inline M3<T> rot_sym(const AxRot<T>& r, const M3<T>& A) {
  int a = r.a, b = r.b, c = r.c;
  T Bbb = r.cs * A.m[b][b] - r.sn * A.m[c][b], Bbc = r.cs * A.m[b][c] - r.sn * A.m[c][c];
  T Bcb = r.sn * A.m[b][b] + r.cs * A.m[c][b], Bcc = r.sn * A.m[b][c] + r.cs * A.m[c][c];
  M3<T> o;
  o.m[a][a] = A.m[a][a];
  o.m[a][b] = o.m[b][a] = r.cs * A.m[a][b] - r.sn * A.m[a][c];
  o.m[a][c] = o.m[c][a] = r.sn * A.m[a][b] + r.cs * A.m[a][c];
  o.m[b][b] = r.cs * Bbb - r.sn * Bbc;
  o.m[b][c] = o.m[c][b] = r.sn * Bbb + r.cs * Bbc;
  o.m[c][c] = r.sn * Bcb + r.cs * Bcc;
  return o;
}
// symmetric helpers: only the upper triangle is computed
template <class T> inline M3<T> add_sym(const M3<T>& A, const M3<T>& B) {
  M3<T> C;
  for (int i = 0; i < 3; i++)
    for (int j = i; j < 3; j++) { C.m[i][j] = A.m[i][j] + B.m[i][j]; C.m[j][i] = C.m[i][j]; }
  return C;
}

// ----------------------------------------------------------------------------------------------
// spatial quantities in link coordinates (Featherstone): motion [w; v], force [n; f]
// ----------------------------------------------------------------------------------------------
template <class T> struct SV { V3<T> a, l; };  // angular, linear
// articulated-body inertia [[I, H], [H^T, M]]
template <class T> struct ABI { M3<T> I, H, M; };

template <class T> inline ABI<T> rigid_inertia(T mass, const V3<T>& com, const T* Ic /*xx yy zz xy xz yz*/) {
  ABI<T> A;
  V3<T> h = mass * com;
  // inertia about the link origin: Ic + m (c.c 1 - c c^T)
  T cc = dot(com, com);
  T I[3][3] = {{Ic[0], Ic[3], Ic[4]}, {Ic[3], Ic[1], Ic[5]}, {Ic[4], Ic[5], Ic[2]}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) A.I.m[i][j] = I[i][j] + mass * ((i == j ? cc : T(0)) - com[i] * com[j]);
  A.H = skew(h);
  A.M.m[0][0] = A.M.m[1][1] = A.M.m[2][2] = mass;
  return A;
}
template <class T> inline SV<T> abi_mul(const ABI<T>& A, const SV<T>& v) {
  SV<T> r;
  r.a = mul(A.I, v.a) + mul(A.H, v.l);
  r.l = mulT(A.H, v.a) + mul(A.M, v.l);
  return r;
}
// v x* f  (spatial force cross product)
template <class T> inline SV<T> crf(const SV<T>& v, const SV<T>& f) {
  SV<T> r;
  r.a = cross(v.a, f.a) + cross(v.l, f.l);
  r.l = cross(v.a, f.l);
  return r;
}

// ----------------------------------------------------------------------------------------------
// model
// ----------------------------------------------------------------------------------------------
template <class T> struct Model {
  T dt, gz, action_scale, action_clip, kn, cn, mu, dtan, radius, veps;
  int nsub;
  T base_inertial[10];
  T lumps[2][10];
  struct Body { T mass; V3<T> com; T Ic[6]; V3<T> r; int axis; } body[12];
  V3<T> foot[4];
  T qdef[12], tlim[12], kp[12], kd[12];

  bool load(const float* b, int n) {
    if (n < SPI_BLOB_SIZE || b[SPI_BLOB_MAGIC] != SPI_BLOB_MAGIC_VALUE) return false;
    dt = T((double)b[SPI_BLOB_DT]); gz = T((double)b[SPI_BLOB_GRAVITY_Z]);
    action_scale = T((double)b[SPI_BLOB_ACTION_SCALE]); action_clip = T((double)b[SPI_BLOB_ACTION_CLIP]);
    kn = T((double)b[SPI_BLOB_CONTACT_KN]); cn = T((double)b[SPI_BLOB_CONTACT_CN]);
    mu = T((double)b[SPI_BLOB_CONTACT_MU]); dtan = T((double)b[SPI_BLOB_CONTACT_DT]);
    radius = T((double)b[SPI_BLOB_FOOT_RADIUS]); veps = T((double)b[SPI_BLOB_CONTACT_VEPS]);
    nsub = (int)b[SPI_BLOB_NSUB]; if (nsub < 1) nsub = 1;
    for (int k = 0; k < 10; k++) base_inertial[k] = T((double)b[SPI_BLOB_BASE_INERTIAL + k]);
    for (int l = 0; l < 2; l++) for (int k = 0; k < 10; k++) lumps[l][k] = T((double)b[SPI_BLOB_BASE_LUMPS + 10 * l + k]);
    for (int i = 0; i < 12; i++) {
      const float* p = b + SPI_BLOB_LEG_BODIES + SPI_LEG_BODY_STRIDE * i;
      body[i].mass = T((double)p[0]);
      body[i].com = V3<T>(T((double)p[1]), T((double)p[2]), T((double)p[3]));
      for (int k = 0; k < 6; k++) body[i].Ic[k] = T((double)p[4 + k]);
      body[i].r = V3<T>(T((double)p[10]), T((double)p[11]), T((double)p[12]));
      body[i].axis = (int)p[13];
    }
    for (int l = 0; l < 4; l++)
      foot[l] = V3<T>(T((double)b[SPI_BLOB_FOOT_OFFSET + 3 * l]), T((double)b[SPI_BLOB_FOOT_OFFSET + 3 * l + 1]),
                      T((double)b[SPI_BLOB_FOOT_OFFSET + 3 * l + 2]));
    for (int j = 0; j < 12; j++) {
      qdef[j] = T((double)b[SPI_BLOB_Q_DEFAULT + j]); tlim[j] = T((double)b[SPI_BLOB_TORQUE_LIMIT + j]);
      kp[j] = T((double)b[SPI_BLOB_KP + j]); kd[j] = T((double)b[SPI_BLOB_KD + j]);
    }
    return true;
  }
};

// Candidate -> base-link inertial record + motor parameters.
// Semantics of the per-env setters of isaacgym_active_sysid.py:61-94 (mass / com* / inertia*), with
// mass_scale as in scripts/mass_opt.py:158-160 (base_nominal * mass_scale).  When the mass changes the
// whole URDF inertia tensor is scaled with it unless SPI_FLAG_INERTIA_KEEP; explicitly
// given inertia entries then overwrite.  SPI_FLAG_STRICT_INERTIAY drops `inertiay` like the reference's
// `set_inertiaiy` typo (isaacgym_active_sysid.py:86).
template <class T> struct Candidate {
  T base[10];
  T motor[3];
};
template <class T>
inline Candidate<T> apply_params(const Model<T>& M, const float* params, int P, const int* ids, unsigned flags) {
  Candidate<T> c;
  for (int k = 0; k < 10; k++) c.base[k] = M.base_inertial[k];
  c.motor[0] = c.motor[1] = c.motor[2] = T(20.0);
  T mass = M.base_inertial[0];
  for (int p = 0; p < P; p++) {
    if (ids[p] == SPI_PARAM_MASS) mass = T((double)params[p]);
    if (ids[p] == SPI_PARAM_MASS_SCALE) mass = M.base_inertial[0] * T((double)params[p]);
  }
  if (!(flags & SPI_FLAG_INERTIA_KEEP)) {
    T s = mass / M.base_inertial[0];
    for (int k = 4; k < 10; k++) c.base[k] = c.base[k] * s;
  }
  c.base[0] = mass;
  for (int p = 0; p < P; p++) {
    T v = T((double)params[p]);
    switch (ids[p]) {
      case SPI_PARAM_COMX: c.base[1] = v; break;
      case SPI_PARAM_COMY: c.base[2] = v; break;
      case SPI_PARAM_COMZ: c.base[3] = v; break;
      case SPI_PARAM_INERTIAX: c.base[4] = v; break;
      case SPI_PARAM_INERTIAY: if (!(flags & SPI_FLAG_STRICT_INERTIAY)) c.base[5] = v; break;
      case SPI_PARAM_INERTIAZ: c.base[6] = v; break;
      case SPI_PARAM_INERTIAXY: c.base[7] = v; break;
      case SPI_PARAM_INERTIAXZ: c.base[8] = v; break;
      case SPI_PARAM_INERTIAYZ: c.base[9] = v; break;
      case SPI_PARAM_MOTOR_HIP: c.motor[0] = v; break;
      case SPI_PARAM_MOTOR_THIGH: c.motor[1] = v; break;
      case SPI_PARAM_MOTOR_CALF: c.motor[2] = v; break;
      default: break;
    }
  }
  return c;
}

// spatial inertia of (candidate base link + fixed head links) about the base-link origin
template <class T> inline ABI<T> base_inertia(const Model<T>& M, const Candidate<T>& c) {
  ABI<T> A = rigid_inertia<T>(c.base[0], V3<T>(c.base[1], c.base[2], c.base[3]), c.base + 4);
  for (int l = 0; l < 2; l++) {
    const T* L = M.lumps[l];
    ABI<T> B = rigid_inertia<T>(L[0], V3<T>(L[1], L[2], L[3]), L + 4);
    A.I = add(A.I, B.I); A.H = add(A.H, B.H); A.M = add(A.M, B.M);
  }
  return A;
}

// ----------------------------------------------------------------------------------------------
// state:  pos[3] quat_xyzw[4] v_world[3] w_world[3] q[12] qd[12]     (SPI_STATE_DIM = 37)
// ----------------------------------------------------------------------------------------------
template <class T> struct State {
  V3<T> p; T quat[4]; V3<T> v, w; T q[12], qd[12];
  void from_floats(const float* s) {
    p = V3<T>(T((double)s[0]), T((double)s[1]), T((double)s[2]));
    for (int k = 0; k < 4; k++) quat[k] = T((double)s[3 + k]);
    v = V3<T>(T((double)s[7]), T((double)s[8]), T((double)s[9]));
    w = V3<T>(T((double)s[10]), T((double)s[11]), T((double)s[12]));
    for (int j = 0; j < 12; j++) { q[j] = T((double)s[13 + j]); qd[j] = T((double)s[25 + j]); }
  }
  template <class F> void to(F* s) const {
    s[0] = (F)to_double(p.x); s[1] = (F)to_double(p.y); s[2] = (F)to_double(p.z);
    for (int k = 0; k < 4; k++) s[3 + k] = (F)to_double(quat[k]);
    s[7] = (F)to_double(v.x); s[8] = (F)to_double(v.y); s[9] = (F)to_double(v.z);
    s[10] = (F)to_double(w.x); s[11] = (F)to_double(w.y); s[12] = (F)to_double(w.z);
    for (int j = 0; j < 12; j++) { s[13 + j] = (F)to_double(q[j]); s[25 + j] = (F)to_double(qd[j]); }
  }
  bool finite() const {
    bool ok = s_finite(p.x) && s_finite(p.y) && s_finite(p.z) && s_finite(v.x) && s_finite(v.y) && s_finite(v.z) &&
              s_finite(w.x) && s_finite(w.y) && s_finite(w.z);
    for (int k = 0; k < 4; k++) ok = ok && s_finite(quat[k]);
    for (int j = 0; j < 12; j++) ok = ok && s_finite(q[j]) && s_finite(qd[j]);
    return ok;
  }
};

// 6x6 symmetric positive-definite solve  A x = b  (LDL^T, no pivoting)
template <class T> inline void solve6(T A[6][6], const T b[6], T x[6]) {
  T L[6][6]; T D[6];
  for (int j = 0; j < 6; j++) {
    T d = A[j][j];
    for (int k = 0; k < j; k++) d = d - L[j][k] * L[j][k] * D[k];
    D[j] = d;
    T inv = T(1) / d;
    for (int i = j + 1; i < 6; i++) {
      T s = A[i][j];
      for (int k = 0; k < j; k++) s = s - L[i][k] * L[j][k] * D[k];
      L[i][j] = s * inv;
    }
  }
  T y[6];
  for (int i = 0; i < 6; i++) { T s = b[i]; for (int k = 0; k < i; k++) s = s - L[i][k] * y[k]; y[i] = s; }
  for (int i = 0; i < 6; i++) y[i] = y[i] / D[i];
  for (int i = 5; i >= 0; i--) { T s = y[i]; for (int k = i + 1; k < 6; k++) s = s - L[k][i] * x[k]; x[i] = s; }
}

// constant per-rollout quantities: link-frame spatial inertias of the 13 bodies
template <class T> struct Bodies {
  ABI<T> I[13];   // 0 = base (candidate + head lumps), 1..12 leg bodies
};
template <class T> inline Bodies<T> make_bodies(const Model<T>& M, const Candidate<T>& cand) {
  Bodies<T> B;
  B.I[0] = base_inertia(M, cand);
  for (int i = 0; i < 12; i++) B.I[i + 1] = rigid_inertia<T>(M.body[i].mass, M.body[i].com, M.body[i].Ic);
  return B;
}

// one component of a cross product
template <class T> inline T cross_comp(const V3<T>& a, const V3<T>& b, int i) {
  int j = (i + 1) % 3, k = (i + 2) % 3;
  return a[j] * b[k] - a[k] * b[j];
}

// Forward dynamics of the 13-body floating-base tree with foot contact
// (Featherstone, RBDA Table 9.4, link coordinates; joint axes are coordinate axes so every joint
// rotation is a planar rotation).
//   tau[12] joint torques; returns base spatial acceleration in base coordinates (gravity included),
//   joint accelerations, and the world-frame foot contact forces.
template <class T> struct Accel { SV<T> a0; T qdd[12]; V3<T> foot_force[4]; M3<T> Rb; V3<T> wb, vb; };

template <class T>
inline Accel<T> forward_dynamics(const Model<T>& M, const Bodies<T>& BI, const State<T>& s, const T* tau,
                                 bool with_contact = true, bool with_gravity = true) {
  Accel<T> out;
  M3<T> Rb = quat_to_matrix(s.quat);
  V3<T> wb = mulT(Rb, s.w), vb = mulT(Rb, s.v);
  out.Rb = Rb; out.wb = wb; out.vb = vb;

  // per-body storage, index 0 = base, 1..12 = leg bodies (hip, thigh, calf per leg)
  SV<T> v[13], c[13], pA[13];
  ABI<T> IA[13];
  AxRot<T> J[13];   // joint rotation (child -> parent)
  M3<T> Rw[13];     // body -> world
  V3<T> pw[13];     // origin in world
  V3<T> U_a[13], U_l[13]; T Dinv[13], u[13];
  int parent[13];

  v[0].a = wb; v[0].l = vb;
  Rw[0] = Rb; pw[0] = s.p;
  IA[0] = BI.I[0];
  {
    SV<T> h = abi_mul(IA[0], v[0]);
    pA[0] = crf(v[0], h);
  }
  for (int i = 1; i <= 12; i++) {
    const auto& B = M.body[i - 1];
    int par = ((i - 1) % 3 == 0) ? 0 : i - 1;
    parent[i] = par;
    int ax = B.axis;
    T qdi = s.qd[i - 1];
    J[i] = make_axrot<T>(ax, s.q[i - 1]);
    v[i].a = rotT(J[i], v[par].a);
    v[i].a[ax] = v[i].a[ax] + qdi;
    v[i].l = rotT(J[i], v[par].l + cross(v[par].a, B.r));
    // c = v x (S qd),  S qd = qd e_ax : (w x e_ax qd, v x e_ax qd)
    int b1 = (ax + 1) % 3, c1 = (ax + 2) % 3;
    c[i].a = V3<T>(); c[i].l = V3<T>();
    c[i].a[b1] = v[i].a[c1] * qdi;  c[i].a[c1] = -(v[i].a[b1] * qdi);
    c[i].l[b1] = v[i].l[c1] * qdi;  c[i].l[c1] = -(v[i].l[b1] * qdi);
    Rw[i] = mul_rot(Rw[par], J[i]);
    pw[i] = pw[par] + mul(Rw[par], B.r);
    IA[i] = BI.I[i];
    SV<T> h = abi_mul(IA[i], v[i]);
    pA[i] = crf(v[i], h);
  }
  // compliant foot contact on the calves
  for (int l = 0; l < 4; l++) {
    out.foot_force[l] = V3<T>();
    if (!with_contact) continue;
    int i = 3 * l + 3;
    V3<T> off = M.foot[l];
    V3<T> pf = pw[i] + mul(Rw[i], off);
    T depth = M.radius - pf.z;
    if (depth > T(0)) {
      V3<T> vf = mul(Rw[i], v[i].l + cross(v[i].a, off));
      T fn = M.kn * depth * (T(1) - M.cn * vf.z);
      fn = s_max(fn, T(0));
      T speed = s_sqrt(vf.x * vf.x + vf.y * vf.y + M.veps * M.veps);
      T coef = s_min(M.dtan, M.mu * fn / speed);
      V3<T> F(-(coef * vf.x), -(coef * vf.y), fn);
      out.foot_force[l] = F;
      V3<T> fc = mulT(Rw[i], F);
      V3<T> nc = cross(off, fc);
      pA[i].a = pA[i].a - nc;
      pA[i].l = pA[i].l - fc;
    }
  }
  // inward pass
  for (int i = 12; i >= 1; i--) {
    const auto& B = M.body[i - 1];
    int ax = B.axis, par = parent[i];
    V3<T> Ua(IA[i].I.m[0][ax], IA[i].I.m[1][ax], IA[i].I.m[2][ax]);   // I  e_ax
    V3<T> Ul(IA[i].H.m[ax][0], IA[i].H.m[ax][1], IA[i].H.m[ax][2]);   // H^T e_ax
    T dinv = T(1) / Ua[ax];
    T ui = tau[i - 1] - pA[i].a[ax];
    U_a[i] = Ua; U_l[i] = Ul; Dinv[i] = dinv; u[i] = ui;
    // Ia = IA - U U^T / D   (I, M symmetric: upper triangle only)
    V3<T> Uad = dinv * Ua, Uld = dinv * Ul;
    ABI<T> Ia;
    for (int r = 0; r < 3; r++) {
      for (int k = r; k < 3; k++) {
        Ia.I.m[r][k] = IA[i].I.m[r][k] - Uad[r] * Ua[k]; Ia.I.m[k][r] = Ia.I.m[r][k];
        Ia.M.m[r][k] = IA[i].M.m[r][k] - Uld[r] * Ul[k]; Ia.M.m[k][r] = Ia.M.m[r][k];
      }
      for (int k = 0; k < 3; k++) Ia.H.m[r][k] = IA[i].H.m[r][k] - Uad[r] * Ul[k];
    }
    // pa = pA + Ia c + U u / D
    SV<T> Ic = abi_mul(Ia, c[i]);
    T ud = ui * dinv;
    SV<T> pa;
    pa.a = pA[i].a + Ic.a + ud * Ua;
    pa.l = pA[i].l + Ic.l + ud * Ul;
    // to parent coordinates: rotate the blocks, then shift the reference point by r
    M3<T> I2 = rot_sym(J[i], Ia.I);
    M3<T> H2 = rot_gen(J[i], Ia.H);
    M3<T> M2 = rot_sym(J[i], Ia.M);
    M3<T> Hp = add(H2, cross_left(B.r, M2));                       // H + rx M
    // I + rx H^T - Hp rx   (symmetric; upper triangle)
    M3<T> Ip;
    for (int r = 0; r < 3; r++)
      for (int k = r; k < 3; k++) {
        V3<T> h2k(H2.m[k][0], H2.m[k][1], H2.m[k][2]);     // row k of H2
        V3<T> hpr(Hp.m[r][0], Hp.m[r][1], Hp.m[r][2]);     // row r of Hp
        Ip.m[r][k] = I2.m[r][k] + cross_comp(B.r, h2k, r) - cross_comp(hpr, B.r, k);
        Ip.m[k][r] = Ip.m[r][k];
      }
    IA[par].I = add_sym(IA[par].I, Ip);
    IA[par].H = add(IA[par].H, Hp);
    IA[par].M = add_sym(IA[par].M, M2);
    V3<T> fl = rot(J[i], pa.l);
    V3<T> fa = rot(J[i], pa.a) + cross(B.r, fl);
    pA[par].a = pA[par].a + fa;
    pA[par].l = pA[par].l + fl;
  }
  // base: a0 = -(IA0)^-1 pA0
  {
    T A[6][6], b[6], x[6];
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < 3; k++) {
        A[r][k] = IA[0].I.m[r][k]; A[r][3 + k] = IA[0].H.m[r][k];
        A[3 + r][k] = IA[0].H.m[k][r]; A[3 + r][3 + k] = IA[0].M.m[r][k];
      }
    for (int k = 0; k < 3; k++) { b[k] = -pA[0].a[k]; b[3 + k] = -pA[0].l[k]; }
    solve6(A, b, x);
    out.a0.a = V3<T>(x[0], x[1], x[2]);
    out.a0.l = V3<T>(x[3], x[4], x[5]);
  }
  // outward pass: joint accelerations
  SV<T> a[13];
  a[0] = out.a0;
  for (int i = 1; i <= 12; i++) {
    const auto& B = M.body[i - 1];
    int par = parent[i], ax = B.axis;
    SV<T> ap;
    ap.a = rotT(J[i], a[par].a) + c[i].a;
    ap.l = rotT(J[i], a[par].l + cross(a[par].a, B.r)) + c[i].l;
    T qdd = (u[i] - (dot(U_a[i], ap.a) + dot(U_l[i], ap.l))) * Dinv[i];
    out.qdd[i - 1] = qdd;
    a[i] = ap;
    a[i].a[ax] = a[i].a[ax] + qdd;
  }
  // uniform gravity: every body accelerates by g, joint accelerations unchanged (RBDA §9.4)
  if (with_gravity) {
    V3<T> g(T(0), T(0), M.gz);
    out.a0.l = out.a0.l + mulT(Rb, g);
  }
  return out;
}

// one integrator sub-step of length h (semi-implicit Euler; quaternion first-order + renormalise)
template <class T>
inline void substep(const Model<T>& M, const Bodies<T>& Ibase, State<T>& s, const T* tau, T h, V3<T>* foot_force = nullptr) {
  Accel<T> A = forward_dynamics(M, Ibase, s, tau);
  if (foot_force) for (int l = 0; l < 4; l++) foot_force[l] = A.foot_force[l];
  for (int j = 0; j < 12; j++) {
    s.qd[j] = s.qd[j] + h * A.qdd[j];
    s.q[j] = s.q[j] + h * s.qd[j];
  }
  // classical acceleration of the base origin = spatial linear part + w x v (body coordinates)
  V3<T> acc_b = A.a0.l + cross(A.wb, A.vb);
  s.w = s.w + h * mul(A.Rb, A.a0.a);
  s.v = s.v + h * mul(A.Rb, acc_b);
  s.p = s.p + h * s.v;
  // q <- normalise(q + h/2 * (w,0) (x) q)
  T hx = T(0.5) * h;
  T x = s.quat[0], y = s.quat[1], z = s.quat[2], w = s.quat[3];
  T nx = x + hx * (s.w.x * w + s.w.y * z - s.w.z * y);
  T ny = y + hx * (s.w.y * w + s.w.z * x - s.w.x * z);
  T nz = z + hx * (s.w.z * w + s.w.x * y - s.w.y * x);
  T nw = w - hx * (s.w.x * x + s.w.y * y + s.w.z * z);
  T inv = T(1) / s_sqrt(nx * nx + ny * ny + nz * nz + nw * nw);
  s.quat[0] = nx * inv; s.quat[1] = ny * inv; s.quat[2] = nz * inv; s.quat[3] = nw * inv;
}

// PD law + clip + motor model: legged_robot_base.py:545,557; go2_omni.py:436-437;
// active_sysid_openloop.py:184-186, 356-400.  `a` is the already clipped action
// (legged_robot_base.py:186-187).
template <class T>
inline void compute_torques(const Model<T>& M, const T* a, const T* q, const T* qd, const T* kp, const T* kd,
                            const T* motor, int motor_model, unsigned flags, T* tau) {
  for (int j = 0; j < 12; j++) {
    T as = a[j] * M.action_scale;
    if ((flags & SPI_FLAG_HIP_HALF) && (j % 3 == 0)) as = as * T(0.5);
    T t = kp[j] * (as + M.qdef[j] - q[j]) - kd[j] * qd[j];
    T g = motor[j % 3];
    if (motor_model == SPI_MOTOR_VEC3_TANH && (flags & SPI_FLAG_TANH_BEFORE_CLIP)) {
      T b = T(1) / g;
      t = g * s_tanh(b * t);
      t = s_clip(t, -M.tlim[j], M.tlim[j]);
    } else {
      t = s_clip(t, -M.tlim[j], M.tlim[j]);
      if (motor_model == SPI_MOTOR_SCALAR) t = t * motor[0];
      else if (motor_model == SPI_MOTOR_VEC3) t = t * g;
      else if (motor_model == SPI_MOTOR_VEC3_TANH) { T b = T(1) / g; t = g * s_tanh(b * t); }
    }
    tau[j] = t;
  }
}

// one physics step (BaseSimulator.simulate_at_each_physics_step): nsub integrator sub-steps under
// constant torques
template <class T>
inline void physics_step(const Model<T>& M, const Bodies<T>& Ibase, State<T>& s, const T* tau, V3<T>* foot_force = nullptr) {
  T h = M.dt / T((double)M.nsub);
  for (int k = 0; k < M.nsub; k++) substep(M, Ibase, s, tau, h, foot_force);
}

// H-step replay of one (candidate, segment): LeggedRobotBase.step x H (legged_robot_base.py:169-209)
// from the recorded initial state (scripts/eval.py:252-267).  out_states: [H,37] or null.
template <class T, class F>
inline void rollout(const Model<T>& M, const Candidate<T>& cand, const float* init, const float* actions,
                    const float* gains, int H, int decimation, int motor_model, unsigned flags, State<T>& s,
                    F* out_states) {
  Bodies<T> Ibase = make_bodies(M, cand);
  s.from_floats(init);
  T kp[12], kd[12];
  for (int j = 0; j < 12; j++) {
    kp[j] = gains ? T((double)gains[j]) : M.kp[j];
    kd[j] = gains ? T((double)gains[12 + j]) : M.kd[j];
  }
  for (int k = 0; k < H; k++) {
    T a[12], tau[12];
    for (int j = 0; j < 12; j++) a[j] = s_clip(T((double)actions[12 * k + j]), -M.action_clip, M.action_clip);
    for (int d = 0; d < decimation; d++) {
      compute_torques(M, a, s.q, s.qd, kp, kd, cand.motor, motor_model, flags, tau);
      physics_step(M, Ibase, s, tau);
    }
    if (out_states) s.to(out_states + (size_t)SPI_STATE_DIM * k);
  }
}

// scripts/eval.py:287-292: L2 errors of the final state against the recorded target
template <class T> inline void segment_errors(const State<T>& s, const float* target, T err[3]) {
  T dp = T(0), dq = T(0), dj = T(0);
  T e;
  e = s.p.x - T((double)target[0]); dp = dp + e * e;
  e = s.p.y - T((double)target[1]); dp = dp + e * e;
  e = s.p.z - T((double)target[2]); dp = dp + e * e;
  for (int k = 0; k < 4; k++) { e = s.quat[k] - T((double)target[3 + k]); dq = dq + e * e; }
  for (int j = 0; j < 12; j++) { e = s.q[j] - T((double)target[7 + j]); dj = dj + e * e; }
  err[0] = s_sqrt(dp); err[1] = s_sqrt(dq); err[2] = s_sqrt(dj);
}

}  // namespace spi_oracle
