// go2_ws.cuh — warp-specialised floating-base articulated-body dynamics for Go2-family quadrupeds.
//
// Mapping: one CTA = 32 rollouts (lanes) x 5 roles (warps):
//     roles 0..3  : leg FL / FR / RL / RR — a 3-joint chain hip(x) - thigh(y) - calf(y), lane-private
//     role  4     : base — sums the four hips' articulated inertias / bias forces, solves the 6x6 base
//                   system once per rollout, integrates the floating base
// The roles exchange 27 + 22 floats per rollout per sub-step through shared memory (two CTA barriers per
// sub-step).  Compared with the leg-per-lane kernel (aba_leg.cuh) nothing is computed redundantly, the
// per-leg constants are warp-uniform (they are read straight from the kernel-parameter constant bank as
// FFMA operands and cost no registers), the calf's constant joint projection is precomputed, and the
// joint-origin sparsity of the Go2 chain (hip r = (x,y,0), thigh r = (0,y,0), calf r = (0,0,z)) is
// compiled in.
//
// Everything here is __host__ __device__ so that tests/host_emulation can run the very same arithmetic
// on the CPU, role by role, against the oracle (tests/test_ws_emulation.py) — the only way to debug
// device code in a container without a GPU.
//
// Same published algorithm as the oracle (Featherstone, RBDA 2008, Table 9.4), written independently.
#pragma once

#include <cmath>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define WS_HD __host__ __device__ __forceinline__
#else
#define WS_HD inline
#endif

#include "../../include/spi_b200.h"

namespace ws {

// symmetric 3x3 stored as (xx, yy, zz, xy, xz, yz)
WS_HD constexpr int sidx(int i, int j) { return (i == j) ? i : ((i + j == 1) ? 3 : ((i + j == 2) ? 4 : 5)); }

// ---- fast scalar helpers -------------------------------------------------------------------------------
// MUFU.RCP: max error 1 ulp (2^-23 relative) — as accurate as any other fp32 operation of the recursion.  (Round 1 added a
// Newton step; its two dependent instructions sat on every serial chain of the kernel — joint projections, the base solve, the
// motor model — and bought nothing measurable: deviation from the fp64 oracle unchanged, tests/tools/dev_accuracy.py.)
WS_HD float rcp_fast(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}
WS_HD float rsqrt_fast(float x) {
#if defined(__CUDA_ARCH__)
  // bare MUFU.RSQ: rsqrtf() wraps it in a denormal-input rescue (a compare and two predicated multiplies) that sits on the
  // contact model's dependency chain in the middle of phase 1; its arguments here are >= veps^2 resp. ~1 (measured: 38.3 -> 37.9 ms)
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / std::sqrt(x);
#endif
}
// (on by default, spi_active_b200/_lib.py)  tanh through one ex2 and one reciprocal: |abs error| <= ~2e-7 (the motor model a * tanh(tau / a) only needs an
// absolute accuracy: 2e-7 * a ~ 4e-6 N m).  tanh(x) = sign(x) (1 - 2 / (exp(2|x|) + 1))
WS_HD float tanh_fast(float x) {
#if defined(__CUDA_ARCH__) && defined(SPI_WS_FAST_TANH)
  const float ax = fminf(fabsf(x), 15.0f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * 2.8853900817779268f));
  const float r = fmaf(-2.0f, rcp_fast(e + 1.0f), 1.0f);
  return copysignf(r, x);
#else
  return tanhf(x);
#endif
}
// the motor model a tanh(t / a) with k2 = 2 log2(e) / a precomputed: 8 dependent instructions, two of them MUFU
constexpr float kTwoLog2e = 2.8853900817779268f;
WS_HD float motor_tanh(float t, float g, float k2) {
#if defined(__CUDA_ARCH__) && defined(SPI_WS_FAST_TANH)
  const float ax = fminf(fabsf(t) * k2, 15.0f * kTwoLog2e);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax));
  const float r = fmaf(-2.0f, rcp_fast(e + 1.0f), 1.0f);
  return copysignf(g * r, t);
#else
  return g * tanhf(t * (k2 * (1.0f / kTwoLog2e)));
#endif
}

WS_HD void sincos_joint(float q, float* s, float* c) {
#if defined(__CUDA_ARCH__)
#if defined(SPI_WS_FAST_SINCOS)
  // MUFU path: reduce to [-pi, pi] (joint angles are bounded, |q| < ~5), abs error <= 2^-21
  // (measured r2: __sincosf without the explicit reduction is 12 instructions shorter per sub-step and 1.2 - 1.9 % SLOWER)
  const float k = rintf(q * 0.15915494309189535f);
  float r = fmaf(k, -6.2831854820251465f, q);
  r = fmaf(k, 1.7484555314695172e-07f, r);
  __sincosf(r, s, c);
#else
  sincosf(q, s, c);
#endif
#else
  *s = std::sin(q); *c = std::cos(q);
#endif
}

// ---- constants -------------------------------------------------------------------------------------------
struct SimK {
  float dt, gz, action_scale, action_clip, kn, cn, mu, dtan, radius, veps2;
  int nsub;
};

// Every member starts on a 16-byte boundary and sizeof(LegK) is a multiple of 16: the leg index is a run-time (warp-uniform)
// value, so these constants reach the FMA pipe through uniform registers, and with provable alignment the compiler fetches
// them four at a time (LDCU.128) instead of one by one — the scalar fetches were 7.6 % of the leg loop's issue slots.
struct alignas(16) LegK {
  float m[3], pad_m;            // body masses (hip, thigh, calf + foot)
  float h[3][4];                // m * com (+ pad)
  float Io[3][8];               // inertia about the link origin, symmetric storage (+ 2 pad)
  float rh[2];                  // hip joint origin in the base frame (x, y);   z == 0
  float rt;                     // thigh joint origin in the hip frame, y;      x == z == 0
  float rc;                     // calf joint origin in the thigh frame, z;     x == y == 0
  float foot[3], pad_f;         // foot sphere centre in the calf frame
  float qdef[3], pad_q, tlim[3], pad_t;
  // the calf is a leaf: its articulated inertia is its rigid inertia, so the joint projection
  // Ia = IA - U U^T / D (axis y) is a per-leg constant
  float cUa[3], cDinv;
  float cUl[3], pad_u;
  float cIbb, cIbc, cIcc, pad_i;   // (b, c) = (z, x) block of the projected rotational inertia
  float cHb[3], pad_hb, cHc[3], pad_hc;     // rows b, c of the projected coupling block
  float cMa[6], pad_ma[2];         // projected linear block, symmetric storage
  float cUadb, cUadc, pad_ud[2];   // U / D of the calf joint (what phase 2 multiplies with): angular b, c and
  float cUld[3], pad_uld;          // linear components
#if defined(SPI_WS_CALF_HARMONIC)
  // thigh articulated inertia BEFORE its own projection = thigh rigid inertia + the calf's constant projected inertia
  // rotated by the calf angle and shifted by rc: every entry is a trigonometric polynomial a + b cos q + c sin q +
  // d cos 2q + e sin 2q of the calf angle; the non-zero coefficients (calf_mask), packed in entry order
  float calfA2[64];
#endif
};
static_assert(sizeof(LegK) % 16 == 0, "LegK must keep 16-byte alignment in the leg array");

struct ModelK {
  SimK sim;
  float base_inertial[10];
  float lumps[2][10];
  LegK leg[4];
  float kp[12], kd[12];
};

// joint-origin sparsity masks (bit i set <=> component i may be non-zero)
constexpr int kMaskHip = 0b011, kMaskThigh = 0b010, kMaskCalf = 0b100;

// the joint-origin vector of joint J of the chain as a 3-array (masked components are never read)
WS_HD void joint_r(const LegK& L, int j, float* r) {
  if (j == 0) { r[0] = L.rh[0]; r[1] = L.rh[1]; r[2] = 0.f; }
  else if (j == 1) { r[0] = 0.f; r[1] = L.rt; r[2] = 0.f; }
  else { r[0] = 0.f; r[1] = 0.f; r[2] = L.rc; }
}

// ---- small vector helpers ------------------------------------------------------------------------------
WS_HD void cross3(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
WS_HD void sym_mulv(const float* S, const float* v, float* o) {
  o[0] = S[0] * v[0] + S[3] * v[1] + S[4] * v[2];
  o[1] = S[3] * v[0] + S[1] * v[1] + S[5] * v[2];
  o[2] = S[4] * v[0] + S[5] * v[1] + S[2] * v[2];
}
// o += a x b
WS_HD void add_cross(const float* a, const float* b, float* o) {
  o[0] = fmaf(a[1], b[2], fmaf(-a[2], b[1], o[0]));
  o[1] = fmaf(a[2], b[0], fmaf(-a[0], b[2], o[1]));
  o[2] = fmaf(a[0], b[1], fmaf(-a[1], b[0], o[2]));
}
// o += a x r   with r sparse
template <int MASK> WS_HD void add_cross_ar(const float* a, const float* r, float* o) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    if ((MASK >> k) & 1) o[i] = fmaf(a[j], r[k], o[i]);
    if ((MASK >> j) & 1) o[i] = fmaf(-a[k], r[j], o[i]);
  }
}
// o += r x f   with r sparse
template <int MASK> WS_HD void add_cross_rf(const float* r, const float* f, float* o) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    if ((MASK >> j) & 1) o[i] = fmaf(r[j], f[k], o[i]);
    if ((MASK >> k) & 1) o[i] = fmaf(-r[k], f[j], o[i]);
  }
}

// planar rotation helpers for a joint about coordinate axis AX with (a, b, c) cyclic:
//   R e_a = e_a, R e_b = cs e_b + sn e_c, R e_c = -sn e_b + cs e_c      (child -> parent)
template <int AX> struct Ax { static constexpr int a = AX, b = (AX + 1) % 3, c = (AX + 2) % 3; };
template <int AX> WS_HD void rot_up(float cs, float sn, const float* v, float* o) {  // R v
  o[Ax<AX>::a] = v[Ax<AX>::a];
  o[Ax<AX>::b] = cs * v[Ax<AX>::b] - sn * v[Ax<AX>::c];
  o[Ax<AX>::c] = sn * v[Ax<AX>::b] + cs * v[Ax<AX>::c];
}
template <int AX> WS_HD void rot_down(float cs, float sn, const float* v, float* o) {  // R^T v
  o[Ax<AX>::a] = v[Ax<AX>::a];
  o[Ax<AX>::b] = cs * v[Ax<AX>::b] + sn * v[Ax<AX>::c];
  o[Ax<AX>::c] = cs * v[Ax<AX>::c] - sn * v[Ax<AX>::b];
}

struct Twist { float a[3]; float l[3]; };   // motion [w; v] or force [n; f]

// what a joint keeps between the inward and the acceleration pass
// (U / D and u / D rather than U, 1 / D and u: phase 2 is then  qdd = u/D - (U/D) . a'  without a final multiply, and the joint
// keeps 12 floats instead of 14; (U/D)[a] == 1 is not stored)
struct Keep {
  float cs, sn;
  float cab, cac, clb, clc;  // velocity-product acceleration c = v x (e_a qd): components b, c
  float Uadb, Uadc, Uld[3], ud;
};

// ---- outward pass for one joint: velocity, velocity-product terms, bias force of the rigid body ---------
template <int AX, int MASK>
WS_HD void joint_outward(const Twist& vp, const float* r, float qd, float mass, const float* h,
                         const float* Io, Twist& v, Keep& k, Twist& pA) {
  constexpr int a = Ax<AX>::a, b = Ax<AX>::b, c = Ax<AX>::c;
  // (k.cs / k.sn of the joint angle are set by leg_angles)
  float t[3] = {vp.l[0], vp.l[1], vp.l[2]};
  add_cross_ar<MASK>(vp.a, r, t);
  rot_down<AX>(k.cs, k.sn, vp.a, v.a);
  v.a[a] += qd;
  rot_down<AX>(k.cs, k.sn, t, v.l);
  k.cab = v.a[c] * qd;  k.cac = -(v.a[b] * qd);
  k.clb = v.l[c] * qd;  k.clc = -(v.l[b] * qd);
  // momentum: n = Io w + h x v,  f = m v - h x w ;  pA = [w x n + v x f ; w x f]
  // (cross products that are only ever added to something are accumulated as FMA chains: the kernel is bound by issue
  // slots, and a separate FMUL / FADD per component costs a slot each)
  float n[3], f[3], hw[3];
  sym_mulv(Io, v.a, n);
  add_cross(h, v.l, n);
  cross3(h, v.a, hw);
#pragma unroll
  for (int i = 0; i < 3; i++) f[i] = mass * v.l[i] - hw[i];
  cross3(v.a, n, pA.a);
  add_cross(v.l, f, pA.a);
  cross3(v.a, f, pA.l);
}

// articulated inertia [[I, H], [H^T, M]], I and M symmetric
struct ABI { float I[6]; float H[9]; float M[6]; };

WS_HD void abi_from_rigid(float mass, const float* h, const float* Io, ABI& A) {
#pragma unroll
  for (int i = 0; i < 6; i++) A.I[i] = Io[i];
  A.H[0] = 0.f;   A.H[1] = -h[2]; A.H[2] = h[1];
  A.H[3] = h[2];  A.H[4] = 0.f;   A.H[5] = -h[0];
  A.H[6] = -h[1]; A.H[7] = h[0];  A.H[8] = 0.f;
  A.M[0] = A.M[1] = A.M[2] = mass;
  A.M[3] = A.M[4] = A.M[5] = 0.f;
}

// The projected quantities of a joint (row/column a of I and row a of H vanish identically).
struct Proj { float Ibb, Ibc, Icc, Hb[3], Hc[3], Ma[6]; };

// ---- force of a joint's subtree to the parent: the rotated linear force is needed by itself (r x f), the torque only as a sum
template <int AX, int MASK, bool PARENT_ZERO = false>
WS_HD void force_to_parent(const float* pa_a, const float* pa_l, float cs, float sn, const float* r, Twist& pAp) {
  constexpr int a = Ax<AX>::a, b = Ax<AX>::b, c = Ax<AX>::c;
  float fl[3];
  rot_up<AX>(cs, sn, pa_l, fl);
  if (PARENT_ZERO) {
    rot_up<AX>(cs, sn, pa_a, pAp.a);
#pragma unroll
    for (int i = 0; i < 3; i++) pAp.l[i] = fl[i];
  } else {
    pAp.a[a] += pa_a[a];
    pAp.a[b] = fmaf(cs, pa_a[b], fmaf(-sn, pa_a[c], pAp.a[b]));
    pAp.a[c] = fmaf(sn, pa_a[b], fmaf(cs, pa_a[c], pAp.a[c]));
#pragma unroll
    for (int i = 0; i < 3; i++) pAp.l[i] += fl[i];
  }
  add_cross_rf<MASK>(r, fl, pAp.a);
}

// ---- transform a projected inertia + force to the parent and accumulate -------------------------------
// pa = pA + Ia c + U u / D must be given; IAp / pAp already hold the parent's own inertia / bias force.
// PARENT_ZERO: IAp / pAp hold nothing yet (the hip's parent is the base, whose own inertia the base role adds) — the results
// are assigned instead of added to zeros.  PARENT_RIGID: IAp was filled by abi_from_rigid, i.e. M is diagonal and H is skew —
// the structurally zero entries are assigned as well (x + 0.0f is not folded by the compiler: -0 + 0 = +0).
// AD ("axis-decoupled"): the projected linear block has M[a][b] = M[a][c] = 0 — true for the thigh of a Go2-family leg: its
// own M is diagonal, and the calf's share has no such entries because the coupling block of a rigid LEAF is skew, so U_l of the
// calf joint has no component along the (parallel) joint axis — a structural zero for any rigid calf, not a symmetry of the
// Go2's (tests/test_ws_emulation.py: an asymmetric calf).  The products with those zeros are skipped (x * 0.0f is not folded).
template <int AX, int MASK, bool PARENT_ZERO = false, bool PARENT_RIGID = false, bool AD = false>
WS_HD void project_to_parent(const Proj& P, const float* pa_a, const float* pa_l, float cs, float sn, const float* r,
                             ABI& IAp, Twist& pAp) {
  constexpr int a = Ax<AX>::a, b = Ax<AX>::b, c = Ax<AX>::c;
  auto mzero = [](int i, int j) constexpr { return AD && ((i == a) != (j == a)); };   // M2[i][j] is structurally zero
  // rotate the blocks to parent orientation:  X' = R X R^T
  //   I: only the (b,c) 2x2 block is non-zero; its second rotation is folded into the accumulation of Ip below
  const float t1 = cs * P.Ibb - sn * P.Ibc, t2 = cs * P.Ibc - sn * P.Icc;
  const float t3 = sn * P.Ibb + cs * P.Ibc, t4 = sn * P.Ibc + cs * P.Icc;
  //   H: rows b, c non-zero
  float Xb[3], Xc[3], H2b[3], H2c[3];
#pragma unroll
  for (int j = 0; j < 3; j++) { Xb[j] = cs * P.Hb[j] - sn * P.Hc[j]; Xc[j] = sn * P.Hb[j] + cs * P.Hc[j]; }
  H2b[a] = Xb[a]; H2b[b] = cs * Xb[b] - sn * Xb[c]; H2b[c] = sn * Xb[b] + cs * Xb[c];
  H2c[a] = Xc[a]; H2c[b] = cs * Xc[b] - sn * Xc[c]; H2c[c] = sn * Xc[b] + cs * Xc[c];
  //   M: full symmetric
  float M2[6];
  {
    const float Bbb = cs * P.Ma[sidx(b, b)] - sn * P.Ma[sidx(c, b)], Bbc = cs * P.Ma[sidx(b, c)] - sn * P.Ma[sidx(c, c)];
    const float Bcb = sn * P.Ma[sidx(b, b)] + cs * P.Ma[sidx(c, b)], Bcc = sn * P.Ma[sidx(b, c)] + cs * P.Ma[sidx(c, c)];
    M2[sidx(a, a)] = P.Ma[sidx(a, a)];
    M2[sidx(a, b)] = AD ? 0.f : cs * P.Ma[sidx(a, b)] - sn * P.Ma[sidx(a, c)];
    M2[sidx(a, c)] = AD ? 0.f : sn * P.Ma[sidx(a, b)] + cs * P.Ma[sidx(a, c)];
    M2[sidx(b, b)] = cs * Bbb - sn * Bbc;
    M2[sidx(b, c)] = sn * Bbb + cs * Bbc;
    M2[sidx(c, c)] = sn * Bcb + cs * Bcc;
  }
  // shift the reference point by r:  Hp = H2 + r x M2 ;  Ip = I2 + r x H2^T - Hp r x
  // (H2 row a is identically zero; masked components of r are skipped at compile time)
  float Hp[9];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      float v = (i == b) ? H2b[j] : ((i == c) ? H2c[j] : 0.f);
      bool have = (i != a);
      if (((MASK >> i1) & 1) && !mzero(i2, j)) { v = have ? fmaf(r[i1], M2[sidx(i2, j)], v) : r[i1] * M2[sidx(i2, j)]; have = true; }
      if (((MASK >> i2) & 1) && !mzero(i1, j)) { v = have ? fmaf(-r[i2], M2[sidx(i1, j)], v) : -(r[i2] * M2[sidx(i1, j)]); have = true; }
      Hp[3 * i + j] = have ? v : 0.f;
    }
  }
  // rows of Hp that are structurally zero: row i is zero iff i == a and neither r[i1] nor r[i2] is present
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = i; j < 3; j++) {
      // Ip[i][j] = parent + I2[i][j] + (r x H2row_j)_i - (Hprow_i x r)_j, one FMA chain starting from the parent's entry
      // (I2bb = t1 cs - t2 sn, I2bc = t1 sn + t2 cs, I2cc = t3 sn + t4 cs)
      float v = PARENT_ZERO ? 0.f : IAp.I[sidx(i, j)];
      bool have = !PARENT_ZERO;
      if (i != a && j != a) {
        const float x1 = (i == c && j == c) ? t3 : t1, x2 = (i == c && j == c) ? t4 : t2;
        const float y1 = (i == b && j == b) ? cs : sn, y2 = (i == b && j == b) ? -sn : cs;     // bb: cs, -sn; bc / cc: sn, cs
        v = have ? fmaf(x1, y1, fmaf(x2, y2, v)) : fmaf(x1, y1, x2 * y2);
        have = true;
      }
      // (r x x)_i = r[i1] x[i2] - r[i2] x[i1],  x = row j of H2 (zero when j == a)
      if (j != a) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        const float* x = (j == b) ? H2b : H2c;
        if ((MASK >> i1) & 1) { v = have ? fmaf(r[i1], x[i2], v) : r[i1] * x[i2]; have = true; }
        if ((MASK >> i2) & 1) { v = have ? fmaf(-r[i2], x[i1], v) : -(r[i2] * x[i1]); have = true; }
      }
      // (y x r)_j = y[j1] r[j2] - y[j2] r[j1],  y = row i of Hp
      {
        const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        const bool row_zero = (i == a) && !((MASK >> ((i + 1) % 3)) & 1) && !((MASK >> ((i + 2) % 3)) & 1);
        if (!row_zero) {
          if ((MASK >> j2) & 1) { v = have ? fmaf(-Hp[3 * i + j1], r[j2], v) : -(Hp[3 * i + j1] * r[j2]); have = true; }
          if ((MASK >> j1) & 1) { v = have ? fmaf(Hp[3 * i + j2], r[j1], v) : Hp[3 * i + j2] * r[j1]; have = true; }
        }
      }
      if (have) IAp.I[sidx(i, j)] = v;
      else if (PARENT_ZERO) IAp.I[sidx(i, j)] = 0.f;
    }
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const bool row_zero = (i == a) && !((MASK >> ((i + 1) % 3)) & 1) && !((MASK >> ((i + 2) % 3)) & 1);
#pragma unroll
    for (int j = 0; j < 3; j++) {
      if (PARENT_ZERO) IAp.H[3 * i + j] = row_zero ? 0.f : Hp[3 * i + j];
      else if (!row_zero) IAp.H[3 * i + j] = (PARENT_RIGID && i == j) ? Hp[3 * i + j] : IAp.H[3 * i + j] + Hp[3 * i + j];
    }
  }
#pragma unroll
  for (int i = 0; i < 6; i++) IAp.M[i] = (PARENT_ZERO || (PARENT_RIGID && i >= 3)) ? M2[i] : IAp.M[i] + M2[i];
  force_to_parent<AX, MASK, PARENT_ZERO>(pa_a, pa_l, cs, sn, r, pAp);
}

// ---- inward pass for a joint with a state-dependent articulated inertia (hip, thigh) ------------------------
// AD: IA has H[a][a] = 0 and M[a][b] = M[a][c] = 0 (see project_to_parent); then U_l[a] = 0 and the a-row / a-column of the
// projected linear block is the one of IA.
template <int AX, int MASK, bool PARENT_ZERO = false, bool PARENT_RIGID = false, bool AD = false>
WS_HD void joint_inward(const ABI& IA, const Twist& pA, float tau, const float* r, Keep& k, ABI& IAp, Twist& pAp) {
  constexpr int a = Ax<AX>::a, b = Ax<AX>::b, c = Ax<AX>::c;
  float Ua[3], Ul[3];
#pragma unroll
  for (int i = 0; i < 3; i++) { Ua[i] = IA.I[sidx(i, a)]; Ul[i] = IA.H[3 * a + i]; }
  const float dinv = rcp_fast(Ua[a]);
  const float u = tau - pA.a[a];
  k.Uadb = Ua[b] * dinv; k.Uadc = Ua[c] * dinv;
#pragma unroll
  for (int i = 0; i < 3; i++) k.Uld[i] = (AD && i == a) ? 0.f : Ul[i] * dinv;
  Proj P;
  P.Ibb = IA.I[sidx(b, b)] - k.Uadb * Ua[b];
  P.Ibc = IA.I[sidx(b, c)] - k.Uadb * Ua[c];
  P.Icc = IA.I[sidx(c, c)] - k.Uadc * Ua[c];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    P.Hb[j] = (AD && j == a) ? IA.H[3 * b + j] : IA.H[3 * b + j] - k.Uadb * Ul[j];
    P.Hc[j] = (AD && j == a) ? IA.H[3 * c + j] : IA.H[3 * c + j] - k.Uadc * Ul[j];
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = i; j < 3; j++) {
      if (AD && (i == a || j == a)) P.Ma[sidx(i, j)] = (i == j) ? IA.M[sidx(i, j)] : 0.f;
      else P.Ma[sidx(i, j)] = IA.M[sidx(i, j)] - k.Uld[i] * Ul[j];
    }
  // pa = pA + Ia c + U u / D     (c has no component along a;  pA[a] + U[a] u / D = pA[a] + u = tau)
  const float ud = u * dinv;
  k.ud = ud;
  float pa_a[3], pa_l[3];
  pa_a[a] = tau;
  pa_a[b] = pA.a[b] + P.Ibb * k.cab + P.Ibc * k.cac + P.Hb[b] * k.clb + P.Hb[c] * k.clc + ud * Ua[b];
  pa_a[c] = pA.a[c] + P.Ibc * k.cab + P.Icc * k.cac + P.Hc[b] * k.clb + P.Hc[c] * k.clc + ud * Ua[c];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    if (AD && j == a) pa_l[j] = pA.l[j] + P.Hb[j] * k.cab + P.Hc[j] * k.cac;
    else pa_l[j] = pA.l[j] + P.Hb[j] * k.cab + P.Hc[j] * k.cac + P.Ma[sidx(j, b)] * k.clb + P.Ma[sidx(j, c)] * k.clc + ud * Ul[j];
  }
  project_to_parent<AX, MASK, PARENT_ZERO, PARENT_RIGID, AD>(P, pa_a, pa_l, k.cs, k.sn, r, IAp, pAp);
}

// ---- inward pass for the calf (leaf, axis y): the projection is the per-leg constant in LegK -----------------
WS_HD void calf_inward(const LegK& L, const Twist& pA, float tau, Keep& k, ABI& IAp, Twist& pAp) {
  constexpr int a = 1, b = 2, c = 0;
  const float ud = (tau - pA.a[a]) * L.cDinv;
  k.ud = ud;
  float pa_a[3], pa_l[3];
  pa_a[a] = tau;
  pa_a[b] = pA.a[b] + L.cIbb * k.cab + L.cIbc * k.cac + L.cHb[b] * k.clb + L.cHb[c] * k.clc + ud * L.cUa[b];
  pa_a[c] = pA.a[c] + L.cIbc * k.cab + L.cIcc * k.cac + L.cHc[b] * k.clb + L.cHc[c] * k.clc + ud * L.cUa[c];
#pragma unroll
  for (int j = 0; j < 3; j++)
    pa_l[j] = pA.l[j] + L.cHb[j] * k.cab + L.cHc[j] * k.cac + L.cMa[sidx(j, b)] * k.clb + L.cMa[sidx(j, c)] * k.clc +
              ud * L.cUl[j];
  Proj P;
  P.Ibb = L.cIbb; P.Ibc = L.cIbc; P.Icc = L.cIcc;
#pragma unroll
  for (int j = 0; j < 3; j++) { P.Hb[j] = L.cHb[j]; P.Hc[j] = L.cHc[j]; }
#pragma unroll
  for (int i = 0; i < 6; i++) P.Ma[i] = L.cMa[i];
  const float r[3] = {0.f, 0.f, L.rc};
  project_to_parent<1, kMaskCalf, false, true>(P, pa_a, pa_l, k.cs, k.sn, r, IAp, pAp);   // IAp = the thigh's rigid inertia
}


#if defined(SPI_WS_CALF_HARMONIC)
// ---- the calf's share of the thigh's articulated inertia as harmonics of the calf angle ------------------------------------
// The calf is a leaf, so its projected inertia is constant in the calf frame; seen from the thigh frame (rotation about y by
// the calf angle q, shift by (0, 0, rc)) every entry of  thigh rigid inertia + X(q)^T Ia X(q)  is
//     a + b cos q + c sin q + d cos 2q + e sin 2q
// with constant coefficients: 38 FFMA replace the rotation / shift of the 21 entries.  calf_mask = which coefficients are
// non-zero for the Go2-family chain (bit 0 = a ... bit 4 = e; entry order I (xx, yy, zz, xy, xz, yz), H row-major, M (xx, yy,
// zz, xy, xz, yz)); model_from_blob fits the coefficients by a discrete Fourier sum over the direct evaluation (calf_inward)
// and refuses the fast path if a masked-out coefficient is not negligible (it cannot be for a rigid calf behind a y-axis joint
// with a z-only offset — the check guards the derivation, not the model).  `coef` = the leg's 55 packed coefficients: the
// kernel keeps them in SHARED memory (the leg index is a run-time value, so in the constant bank every coefficient would be
// its own LDCU — that is what made this formulation slower in the first r2 session; from shared memory they arrive four at
// a time).
constexpr int kCalfEntries = 21;
WS_HD constexpr unsigned calf_mask(int e) {
  constexpr unsigned m[kCalfEntries] = {31, 25, 25, 25, 31, 25,  25, 7, 25, 25, 0, 25, 25, 7, 25,  25, 1, 25, 0, 24, 0};
  return m[e];
}
WS_HD constexpr int calf_offset(int e, int bit) {   // index of coefficient (e, bit) in the packed array
  int n = 0;
  for (int i = 0; i < e; i++)
    for (int b = 0; b < 5; b++) n += (calf_mask(i) >> b) & 1;
  for (int b = 0; b < bit; b++) n += (calf_mask(e) >> b) & 1;
  return n;
}
constexpr int kCalfCoefs = calf_offset(kCalfEntries, 0);
static_assert(kCalfCoefs <= 64, "LegK::calfA2 too small");

WS_HD float& abi_entry(ABI& A, int e) { return e < 6 ? A.I[e] : (e < 15 ? A.H[e - 6] : A.M[e - 15]); }

template <int E> WS_HD float calf_entry(const float* coef, const float* basis) {
  float v = 0.f;
  bool have = false;
#pragma unroll
  for (int b = 0; b < 5; b++) {
    if ((calf_mask(E) >> b) & 1) {
      const float k = coef[calf_offset(E, b)];
      if (b == 0) v = k;
      else v = have ? fmaf(k, basis[b], v) : k * basis[b];
      have = true;
    }
  }
  return v;    // entries without any coefficient are the literal 0.f
}
template <int E> WS_HD void calf_entries(const float* coef, const float* basis, ABI& A2) {
  abi_entry(A2, E) = calf_entry<E>(coef, basis);
  if constexpr (E + 1 < kCalfEntries) calf_entries<E + 1>(coef, basis, A2);
}
WS_HD void calf_inertia_harmonic(const float* coef, float cs, float sn, ABI& A2) {
  const float basis[5] = {1.f, cs, sn, cs * cs - sn * sn, 2.f * cs * sn};
  calf_entries<0>(coef, basis, A2);
}
#endif   // SPI_WS_CALF_HARMONIC

// the force half of calf_inward: pa = pA + Ia c + U u / D, rotated / shifted to the thigh and added to its bias force
WS_HD void calf_force(const LegK& L, const Twist& pA, float tau, Keep& k, Twist& pAp) {
  constexpr int a = 1, b = 2, c = 0;
  const float ud = (tau - pA.a[a]) * L.cDinv;
  k.ud = ud;
  float pa_a[3], pa_l[3];
  pa_a[a] = tau;
  pa_a[b] = pA.a[b] + L.cIbb * k.cab + L.cIbc * k.cac + L.cHb[b] * k.clb + L.cHb[c] * k.clc + ud * L.cUa[b];
  pa_a[c] = pA.a[c] + L.cIbc * k.cab + L.cIcc * k.cac + L.cHc[b] * k.clb + L.cHc[c] * k.clc + ud * L.cUa[c];
#pragma unroll
  for (int j = 0; j < 3; j++)
    pa_l[j] = pA.l[j] + L.cHb[j] * k.cab + L.cHc[j] * k.cac + L.cMa[sidx(j, b)] * k.clb + L.cMa[sidx(j, c)] * k.clc +
              ud * L.cUl[j];
  const float r[3] = {0.f, 0.f, L.rc};
  force_to_parent<1, kMaskCalf>(pa_a, pa_l, k.cs, k.sn, r, pAp);
}

// ---- outward acceleration pass for one joint --------------------------------------------------------------
template <int AX, int MASK, bool AD = false>
WS_HD float joint_accel(const Twist& ap, const float* r, float Uadb, float Uadc, const float* Uld, const Keep& k, Twist& acc) {
  constexpr int a = Ax<AX>::a, b = Ax<AX>::b, c = Ax<AX>::c;
  float t[3] = {ap.l[0], ap.l[1], ap.l[2]};
  add_cross_ar<MASK>(ap.a, r, t);
  // R^T a_parent + c as FMA chains that start from the velocity-product term (a separate FADD per component costs an issue slot)
  acc.a[a] = ap.a[a];
  acc.a[b] = fmaf(k.cs, ap.a[b], fmaf(k.sn, ap.a[c], k.cab));
  acc.a[c] = fmaf(k.cs, ap.a[c], fmaf(-k.sn, ap.a[b], k.cac));
  acc.l[a] = t[a];
  acc.l[b] = fmaf(k.cs, t[b], fmaf(k.sn, t[c], k.clb));
  acc.l[c] = fmaf(k.cs, t[c], fmaf(-k.sn, t[b], k.clc));
  // qdd = (u - U . a') / D accumulated onto u / D;  (U / D)[a] == 1
  float e = k.ud - acc.a[a];
  e = fmaf(-Uadb, acc.a[b], e);
  e = fmaf(-Uadc, acc.a[c], e);
#pragma unroll
  for (int i = 0; i < 3; i++)
    if (!(AD && i == a)) e = fmaf(-Uld[i], acc.l[i], e);
  acc.a[a] += e;
  return e;
}

// number of floats a leg hands to the base role: I(6) H(9) M(6) pA(6)
constexpr int kLegOut = 27;
// number of floats the base role hands to the legs: a0(6) R(9) v0(6) pz(1)
constexpr int kBaseOut = 22;
constexpr int kBcA0 = 0, kBcR = 6, kBcV0 = 15, kBcPz = 21;

struct LegState { float q[3], qd[3]; };
struct LegKeep { Keep k1, k2, k3; };

// sin / cos of the joint angles: the only part of phase 1 that does not need the base state of the sub-step — the kernel
// evaluates it (and the torques of a new physics step) BEFORE it waits for that state
WS_HD void leg_angles(const float* q, LegKeep& K) {
  sincos_joint(q[0], &K.k1.sn, &K.k1.cs);
  sincos_joint(q[1], &K.k2.sn, &K.k2.cs);
  sincos_joint(q[2], &K.k3.sn, &K.k3.cs);
}
// ---- leg role, phase 1: outward pass, foot contact, inward pass -> 27 floats for the base ------------------
// Rv0[22] = the base broadcast (a0 unused here).  out[27] = hip-projected inertia/force in base coordinates.
// foot_force (optional): world-frame contact force on this leg's foot.  K.k*.cs / sn must hold sin / cos of the joint angles.
// calf_coef (SPI_WS_CALF_HARMONIC builds): the leg's packed harmonic coefficients (LegK::calfA2, or the kernel's shared-memory copy)
WS_HD void leg_phase1_core(const SimK& S, const LegK& L, const float* bc, const LegState& s, const float* tau, LegKeep& K,
                           float* out, float* foot_force, const float* calf_coef = nullptr) {
  const float* R = bc + kBcR;
  Twist v0;
#pragma unroll
  for (int i = 0; i < 3; i++) { v0.a[i] = bc[kBcV0 + i]; v0.l[i] = bc[kBcV0 + 3 + i]; }
  float r0[3], r1[3], r2[3];
  joint_r(L, 0, r0); joint_r(L, 1, r1); joint_r(L, 2, r2);
  Twist v1, v2, v3, p1, p2, p3;
  joint_outward<0, kMaskHip>(v0, r0, s.qd[0], L.m[0], L.h[0], L.Io[0], v1, K.k1, p1);
  joint_outward<1, kMaskThigh>(v1, r1, s.qd[1], L.m[1], L.h[1], L.Io[1], v2, K.k2, p2);
  joint_outward<1, kMaskCalf>(v2, r2, s.qd[2], L.m[2], L.h[2], L.Io[2], v3, K.k3, p3);
  // foot contact (compliant sphere on the plane z = 0)
  {
    float t3[3], t2[3], t1[3], u3[3], u2[3];
    // (axis y: a, b, c = y, z, x; axis x: a, b, c = x, y, z) — the joint-origin offsets start the FMA chains they are added to
    t3[0] = K.k3.sn * L.foot[2] + K.k3.cs * L.foot[0];
    u3[0] = t3[0]; u3[1] = L.foot[1]; u3[2] = fmaf(K.k3.cs, L.foot[2], fmaf(-K.k3.sn, L.foot[0], L.rc));
    rot_up<1>(K.k2.cs, K.k2.sn, u3, t2);
    u2[0] = t2[0]; u2[1] = L.rt + t2[1]; u2[2] = t2[2];
    t1[0] = u2[0]; t1[2] = K.k1.sn * u2[1] + K.k1.cs * u2[2];
    const float fbx = L.rh[0] + t1[0], fby = fmaf(K.k1.cs, u2[1], fmaf(-K.k1.sn, u2[2], L.rh[1])), fbz = t1[2];
    const float pz = bc[kBcPz] + R[6] * fbx + R[7] * fby + R[8] * fbz;
    const float depth = S.radius - pz;
    float F[3] = {0.f, 0.f, 0.f};
    if (depth > 0.f) {
      float vc[3] = {v3.l[0], v3.l[1], v3.l[2]}, a2[3], a1[3], vb[3], vw[3];
      add_cross(v3.a, L.foot, vc);
      rot_up<1>(K.k3.cs, K.k3.sn, vc, a2);
      rot_up<1>(K.k2.cs, K.k2.sn, a2, a1);
      rot_up<0>(K.k1.cs, K.k1.sn, a1, vb);
#pragma unroll
      for (int i = 0; i < 3; i++) vw[i] = R[3 * i] * vb[0] + R[3 * i + 1] * vb[1] + R[3 * i + 2] * vb[2];
      float fn = S.kn * depth * (1.f - S.cn * vw[2]);
      fn = fmaxf(fn, 0.f);
      const float speed2 = vw[0] * vw[0] + vw[1] * vw[1] + S.veps2;
      const float coef = fminf(S.dtan, S.mu * fn * rsqrt_fast(speed2));
      F[0] = -(coef * vw[0]); F[1] = -(coef * vw[1]); F[2] = fn;
      float fb[3], g1[3], g2[3], fc[3];
#pragma unroll
      for (int i = 0; i < 3; i++) fb[i] = R[i] * F[0] + R[3 + i] * F[1] + R[6 + i] * F[2];
      rot_down<0>(K.k1.cs, K.k1.sn, fb, g1);
      rot_down<1>(K.k2.cs, K.k2.sn, g1, g2);
      rot_down<1>(K.k3.cs, K.k3.sn, g2, fc);
      add_cross(fc, L.foot, p3.a);                    // p3.a -= foot x fc
#pragma unroll
      for (int i = 0; i < 3; i++) p3.l[i] -= fc[i];
    }
    if (foot_force) { foot_force[0] = F[0]; foot_force[1] = F[1]; foot_force[2] = F[2]; }
  }
  // inward pass up the leg
  ABI A2, A1, A0;
#if defined(SPI_WS_CALF_HARMONIC)
  calf_inertia_harmonic(calf_coef, K.k3.cs, K.k3.sn, A2);
  calf_force(L, p3, tau[2], K.k3, p2);
#else
  abi_from_rigid(L.m[1], L.h[1], L.Io[1], A2);
  calf_inward(L, p3, tau[2], K.k3, A2, p2);
#endif
  abi_from_rigid(L.m[0], L.h[0], L.Io[0], A1);
#if defined(SPI_WS_CALF_HARMONIC)
  joint_inward<1, kMaskThigh, false, true, true>(A2, p2, tau[1], r1, K.k2, A1, p1);       // A1 = the hip's rigid inertia
#else
  joint_inward<1, kMaskThigh, false, true>(A2, p2, tau[1], r1, K.k2, A1, p1);       // A1 = the hip's rigid inertia
#endif
  Twist p0;
  joint_inward<0, kMaskHip, true>(A1, p1, tau[0], r0, K.k1, A0, p0);
#pragma unroll
  for (int i = 0; i < 6; i++) { out[i] = A0.I[i]; out[15 + i] = A0.M[i]; }
#pragma unroll
  for (int i = 0; i < 9; i++) out[6 + i] = A0.H[i];
#pragma unroll
  for (int i = 0; i < 3; i++) { out[21 + i] = p0.a[i]; out[24 + i] = p0.l[i]; }
}

WS_HD void leg_phase1(const SimK& S, const LegK& L, const float* bc, const LegState& s, const float* tau, LegKeep& K,
                      float* out, float* foot_force) {
  leg_angles(s.q, K);
#if defined(SPI_WS_CALF_HARMONIC)
  leg_phase1_core(S, L, bc, s, tau, K, out, foot_force, L.calfA2);
#else
  leg_phase1_core(S, L, bc, s, tau, K, out, foot_force);
#endif
}

// ---- leg role, phase 2: acceleration pass + semi-implicit Euler of the 3 joints ------------------------------
WS_HD void leg_phase2(const LegK& L, const float* bc, const LegKeep& K, LegState& s, float h) {
  Twist a0, a1, a2, a3;
#pragma unroll
  for (int i = 0; i < 3; i++) { a0.a[i] = bc[kBcA0 + i]; a0.l[i] = bc[kBcA0 + 3 + i]; }
  float r0[3], r1[3], r2[3];
  joint_r(L, 0, r0); joint_r(L, 1, r1); joint_r(L, 2, r2);
  const float qdd0 = joint_accel<0, kMaskHip>(a0, r0, K.k1.Uadb, K.k1.Uadc, K.k1.Uld, K.k1, a1);
#if defined(SPI_WS_CALF_HARMONIC)
  const float qdd1 = joint_accel<1, kMaskThigh, true>(a1, r1, K.k2.Uadb, K.k2.Uadc, K.k2.Uld, K.k2, a2);
#else
  const float qdd1 = joint_accel<1, kMaskThigh>(a1, r1, K.k2.Uadb, K.k2.Uadc, K.k2.Uld, K.k2, a2);
#endif
  const float qdd2 = joint_accel<1, kMaskCalf>(a2, r2, L.cUadb, L.cUadc, L.cUld, K.k3, a3);
  s.qd[0] += h * qdd0; s.q[0] += h * s.qd[0];
  s.qd[1] += h * qdd1; s.q[1] += h * s.qd[1];
  s.qd[2] += h * qdd2; s.q[2] += h * s.qd[2];
}

// ---- base role ----------------------------------------------------------------------------------------------------
struct BaseState { float p[3], quat[4], v[3], w[3]; };
struct BaseInertia { float m; float h[3]; float Io[6]; };

WS_HD void add_point_inertia(BaseInertia& B, float mass, const float* c, const float* Ic) {
  const float cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
  B.m += mass;
#pragma unroll
  for (int i = 0; i < 3; i++) B.h[i] += mass * c[i];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = i; j < 3; j++) B.Io[sidx(i, j)] += Ic[sidx(i, j)] + mass * ((i == j ? cc : 0.f) - c[i] * c[j]);
}

// rotation matrix (body -> world) and body-frame twist of the base -> bc[R, v0, pz]
WS_HD void base_publish(const BaseState& s, float* bc) {
  float* R = bc + kBcR;
  const float x = s.quat[0], y = s.quat[1], z = s.quat[2], w = s.quat[3];
  const float xx = 2.f * x * x, yy = 2.f * y * y, zz = 2.f * z * z;
  const float xy = 2.f * x * y, xz = 2.f * x * z, yz = 2.f * y * z;
  const float wx = 2.f * w * x, wy = 2.f * w * y, wz = 2.f * w * z;
  R[0] = 1.f - (yy + zz); R[1] = xy - wz;         R[2] = xz + wy;
  R[3] = xy + wz;         R[4] = 1.f - (xx + zz); R[5] = yz - wx;
  R[6] = xz - wy;         R[7] = yz + wx;         R[8] = 1.f - (xx + yy);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    bc[kBcV0 + i] = R[i] * s.w[0] + R[3 + i] * s.w[1] + R[6 + i] * s.w[2];
    bc[kBcV0 + 3 + i] = R[i] * s.v[0] + R[3 + i] * s.v[1] + R[6 + i] * s.v[2];
  }
  bc[kBcPz] = s.p[2];
}

// 6x6 SPD solve  [[I, H], [H^T, M]] [a_w; a_v] = -[pA.a; pA.l]  by block elimination on the 3x3 blocks:
//     Minv = adj(M) / det M,   Y = H Minv,   S = I - Y H^T (Schur complement, SPD),   a_w = S^-1 (b_w - Y b_v),
//     a_v = Minv (b_v - H^T a_w)
// The base solve is the serial section between the two leg phases (the four leg warps wait for it), so what matters is
// the length of its dependency chain: two reciprocals and wide 3x3 products instead of the six sequential pivots of an
// LDL^T factorisation (measured: the LDL^T version ran at 0.6x the legs' IPC — profiles/README.md).  Both 3x3 blocks
// are well conditioned (M ~ total mass, S ~ the composite rotational inertia), so the adjugate inverses are fp32-safe.
WS_HD void sym3_adjugate(const float* A, float* C, float* det) {   // A, C symmetric (xx, yy, zz, xy, xz, yz)
  C[0] = A[1] * A[2] - A[5] * A[5];
  C[1] = A[0] * A[2] - A[4] * A[4];
  C[2] = A[0] * A[1] - A[3] * A[3];
  C[3] = A[4] * A[5] - A[3] * A[2];
  C[4] = A[3] * A[5] - A[4] * A[1];
  C[5] = A[3] * A[4] - A[0] * A[5];
  *det = A[0] * C[0] + A[3] * C[3] + A[4] * C[4];
}
// The angular part is solved UNNORMALISED (S' = det(M) S, t' = det(M) t, so that a_w = adj(S') t' / det(S')): the reciprocal of
// det(M) is then only needed for a_v at the very end and leaves the dependency chain adj(M) -> S -> adj(S) -> a_w, which carries
// ONE MUFU round trip instead of two in a row (magnitudes: det(M) ~ 3e3, det(S') ~ 1e8 ... 1e10 — far from the fp32 range).
WS_HD void solve_base(const ABI& A, const Twist& pA, Twist& a0) {
  float Cm[6], detm;
  sym3_adjugate(A.M, Cm, &detm);
  const float rm = rcp_fast(detm);
  // Yc = H adj(M)   (Y = Yc / det M)
  float Yc[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      Yc[3 * i + j] = A.H[3 * i] * Cm[sidx(0, j)] + A.H[3 * i + 1] * Cm[sidx(1, j)] + A.H[3 * i + 2] * Cm[sidx(2, j)];
  // S' = det(M) I - Yc H^T = det(M) (I - Y H^T), symmetric
  float S[6];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = i; j < 3; j++) {
      const float yh = Yc[3 * i] * A.H[3 * j] + Yc[3 * i + 1] * A.H[3 * j + 1] + Yc[3 * i + 2] * A.H[3 * j + 2];
      S[sidx(i, j)] = fmaf(detm, A.I[sidx(i, j)], -yh);
    }
  // t' = det(M) (b_w - Y b_v)  with b = -pA:  t' = -det(M) pA.a + Yc pA.l
  float t[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float yb = Yc[3 * i] * pA.l[0] + Yc[3 * i + 1] * pA.l[1] + Yc[3 * i + 2] * pA.l[2];
    t[i] = fmaf(-detm, pA.a[i], yb);
  }
  float Cs[6], dets;
  sym3_adjugate(S, Cs, &dets);
  const float rs = rcp_fast(dets);
#pragma unroll
  for (int i = 0; i < 3; i++)
    a0.a[i] = (Cs[sidx(i, 0)] * t[0] + Cs[sidx(i, 1)] * t[1] + Cs[sidx(i, 2)] * t[2]) * rs;
  // a_v = Minv (b_v - H^T a_w) = -(adj(M) pA.l + Yc^T a_w) / det M   (adj(M) H^T = Yc^T): the first product does not wait for
  // a_w, so only one 3-term chain and a scale follow it
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float w = Cm[sidx(i, 0)] * pA.l[0] + Cm[sidx(i, 1)] * pA.l[1] + Cm[sidx(i, 2)] * pA.l[2];
    w = fmaf(Yc[i], a0.a[0], fmaf(Yc[3 + i], a0.a[1], fmaf(Yc[6 + i], a0.a[2], w)));
    a0.l[i] = -(w * rm);
  }
}

// Base role, split in two so that the legs' acceleration pass overlaps the base integration:
//   base_bias    (off the critical path, right after the state is known): velocity-product bias of the base body
//   base_solve   (critical: between the legs' phase 1 and phase 2): sum of the legs + base -> 6x6 solve -> a0
//   base_advance (overlaps the legs' phase 2): semi-implicit Euler of the base, publish R / v0 / pz of the NEW
//                state for phase 1 of the next sub-step
WS_HD void base_bias(const BaseInertia& B, const float* bc, float* pb /*[6]*/) {
  Twist v0;
#pragma unroll
  for (int i = 0; i < 3; i++) { v0.a[i] = bc[kBcV0 + i]; v0.l[i] = bc[kBcV0 + 3 + i]; }
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
  for (int i = 0; i < 3; i++) { pb[i] = t1[i] + t2[i]; pb[3 + i] = t3[i]; }
}

WS_HD void base_solve(const BaseInertia& B, const float* legsum, const float* pb, float* a0_out /*[6]*/) {
  ABI A0;
  Twist p0;
#pragma unroll
  for (int i = 0; i < 6; i++) { A0.I[i] = legsum[i] + B.Io[i]; A0.M[i] = legsum[15 + i]; }
#pragma unroll
  for (int i = 0; i < 9; i++) A0.H[i] = legsum[6 + i];
  A0.H[1] -= B.h[2]; A0.H[2] += B.h[1];
  A0.H[3] += B.h[2]; A0.H[5] -= B.h[0];
  A0.H[6] -= B.h[1]; A0.H[7] += B.h[0];
  A0.M[0] += B.m; A0.M[1] += B.m; A0.M[2] += B.m;
#pragma unroll
  for (int i = 0; i < 3; i++) { p0.a[i] = legsum[21 + i] + pb[i]; p0.l[i] = legsum[24 + i] + pb[3 + i]; }
  Twist a0;
  solve_base(A0, p0, a0);
#pragma unroll
  for (int i = 0; i < 3; i++) { a0_out[i] = a0.a[i]; a0_out[3 + i] = a0.l[i]; }
}

// bc holds R / v0 of the CURRENT state on entry and of the NEW state on return; bc[a0] is left untouched.
WS_HD void base_advance(const SimK& S, const float* a0v, BaseState& s, float h, float* bc) {
  Twist v0, a0;
  float Rl[9];
#pragma unroll
  for (int i = 0; i < 3; i++) { v0.a[i] = bc[kBcV0 + i]; v0.l[i] = bc[kBcV0 + 3 + i]; a0.a[i] = a0v[i]; a0.l[i] = a0v[3 + i]; }
#pragma unroll
  for (int i = 0; i < 9; i++) Rl[i] = bc[kBcR + i];
  // gravity enters as a uniform acceleration of every body (RBDA 9.4): classical acc of the base origin
  float accb[3], wxv[3];
  cross3(v0.a, v0.l, wxv);
#pragma unroll
  for (int i = 0; i < 3; i++) accb[i] = a0.l[i] + Rl[6 + i] * S.gz + wxv[i];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    s.w[i] += h * (Rl[3 * i] * a0.a[0] + Rl[3 * i + 1] * a0.a[1] + Rl[3 * i + 2] * a0.a[2]);
    s.v[i] += h * (Rl[3 * i] * accb[0] + Rl[3 * i + 1] * accb[1] + Rl[3 * i + 2] * accb[2]);
    s.p[i] += h * s.v[i];
  }
  const float hx = 0.5f * h;
  const float x = s.quat[0], y = s.quat[1], z = s.quat[2], w = s.quat[3];
  const float nx = x + hx * (s.w[0] * w + s.w[1] * z - s.w[2] * y);
  const float ny = y + hx * (s.w[1] * w + s.w[2] * x - s.w[0] * z);
  const float nz = z + hx * (s.w[2] * w + s.w[0] * y - s.w[1] * x);
  const float nw = w - hx * (s.w[0] * x + s.w[1] * y + s.w[2] * z);
  const float inv = rsqrt_fast(nx * nx + ny * ny + nz * nz + nw * nw);
  s.quat[0] = nx * inv; s.quat[1] = ny * inv; s.quat[2] = nz * inv; s.quat[3] = nw * inv;
  base_publish(s, bc);
}

// PD law + torque clip + motor model for one leg's 3 joints
// (legged_robot_base.py:545,557; go2_omni.py:436-437; active_sysid_openloop.py:184-186,356-400)
// The evaluation sits in front of the dynamics of every physics step as a serial section of the leg warp, so it is kept short:
//   * MODEL / TANH_FIRST are compile-time: the three joints are ONE basic block whose MUFU round trips overlap (with run-time
//     branches per joint every joint was its own serial chain: 1 360 cycles per physics step, tools/ws_timeline.py);
//   * what only changes with the control step is precomputed once per control step (leg_pd_bias): kp (a s + q_default);
//   * what never changes is precomputed once per rollout: k2 = 2 log2(e) / motor gain.
WS_HD void leg_pd_bias(const SimK& S, const LegK& L, const float* act /*clipped*/, const float* kp, float hip_scale, float* pdb) {
#pragma unroll
  for (int j = 0; j < 3; j++) {
    float as = act[j] * S.action_scale;
    if (j == 0) as *= hip_scale;
    pdb[j] = kp[j] * (as + L.qdef[j]);
  }
}
template <int MODEL, bool TANH_FIRST>
WS_HD void leg_torques_t(const LegK& L, const float* pdb, const float* q, const float* qd, const float* kp, const float* kd,
                         const float* motor, const float* motor_k2, float* tau) {
#pragma unroll
  for (int j = 0; j < 3; j++) {
    float t = fmaf(-kd[j], qd[j], fmaf(-kp[j], q[j], pdb[j]));      // kp (a s + q_default - q) - kd qd
    const float g = motor[j];
    if (MODEL == SPI_MOTOR_VEC3_TANH && TANH_FIRST) {
      t = motor_tanh(t, g, motor_k2[j]);
      t = fminf(fmaxf(t, -L.tlim[j]), L.tlim[j]);
    } else {
      t = fminf(fmaxf(t, -L.tlim[j]), L.tlim[j]);
      if (MODEL == SPI_MOTOR_SCALAR) t *= motor[0];
      else if (MODEL == SPI_MOTOR_VEC3) t *= g;
      else if (MODEL == SPI_MOTOR_VEC3_TANH) t = motor_tanh(t, g, motor_k2[j]);
    }
    tau[j] = t;
  }
}
// the motor model is dispatched AROUND the joint loop
WS_HD void leg_torques_dispatch(const LegK& L, const float* pdb, const float* q, const float* qd, const float* kp,
                                const float* kd, const float* motor, const float* motor_k2, int motor_model, unsigned flags,
                                float* tau) {
  if (motor_model == SPI_MOTOR_VEC3_TANH) {
    if (flags & SPI_FLAG_TANH_BEFORE_CLIP) leg_torques_t<SPI_MOTOR_VEC3_TANH, true>(L, pdb, q, qd, kp, kd, motor, motor_k2, tau);
    else leg_torques_t<SPI_MOTOR_VEC3_TANH, false>(L, pdb, q, qd, kp, kd, motor, motor_k2, tau);
  } else if (motor_model == SPI_MOTOR_SCALAR) leg_torques_t<SPI_MOTOR_SCALAR, false>(L, pdb, q, qd, kp, kd, motor, motor_k2, tau);
  else if (motor_model == SPI_MOTOR_VEC3) leg_torques_t<SPI_MOTOR_VEC3, false>(L, pdb, q, qd, kp, kd, motor, motor_k2, tau);
  else leg_torques_t<SPI_MOTOR_NONE, false>(L, pdb, q, qd, kp, kd, motor, motor_k2, tau);
}

// everything from the clipped action (host emulation, tests)
WS_HD void leg_torques(const SimK& S, const LegK& L, const float* act /*clipped*/, const float* q, const float* qd,
                       const float* kp, const float* kd, const float* motor, int motor_model, unsigned flags, float* tau) {
  float pdb[3], k2[3];
  leg_pd_bias(S, L, act, kp, (flags & SPI_FLAG_HIP_HALF) ? 0.5f : 1.0f, pdb);
#pragma unroll
  for (int j = 0; j < 3; j++) k2[j] = kTwoLog2e * rcp_fast(motor[j]);
  leg_torques_dispatch(L, pdb, q, qd, kp, kd, motor, k2, motor_model, flags, tau);
}

// candidate row -> base rigid inertia (+ head lumps) and motor parameters
// (isaacgym_active_sysid.py:61-94 setters; mass_opt.py:158-160 mass_scale; DESIGN.md D8/D15 flags)
struct ParamIdsK { int n; int id[16]; };

WS_HD void apply_candidate(const ModelK& M, const float* row, const ParamIdsK& ids, unsigned flags, BaseInertia& B,
                           float* motor3) {
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

// host-side: blob -> ModelK (incl. the constant calf projection).  Returns 0, or a negative code when the
// blob does not have the Go2-family structure this fast path is compiled for (the caller then uses the
// generic leg-per-lane kernel): -1 axes, -2 joint-origin sparsity, -3 (SPI_WS_CALF_HARMONIC builds) harmonic structure of the calf's inertia.
inline int model_from_blob(const float* b, ModelK* M) {
  M->sim.dt = b[SPI_BLOB_DT]; M->sim.gz = b[SPI_BLOB_GRAVITY_Z];
  M->sim.action_scale = b[SPI_BLOB_ACTION_SCALE]; M->sim.action_clip = b[SPI_BLOB_ACTION_CLIP];
  M->sim.kn = b[SPI_BLOB_CONTACT_KN]; M->sim.cn = b[SPI_BLOB_CONTACT_CN]; M->sim.mu = b[SPI_BLOB_CONTACT_MU];
  M->sim.dtan = b[SPI_BLOB_CONTACT_DT]; M->sim.radius = b[SPI_BLOB_FOOT_RADIUS];
  M->sim.veps2 = b[SPI_BLOB_CONTACT_VEPS] * b[SPI_BLOB_CONTACT_VEPS];
  M->sim.nsub = (int)b[SPI_BLOB_NSUB];
  if (M->sim.nsub < 1) M->sim.nsub = 1;
  for (int k = 0; k < 10; k++) M->base_inertial[k] = b[SPI_BLOB_BASE_INERTIAL + k];
  for (int l = 0; l < 2; l++)
    for (int k = 0; k < 10; k++) M->lumps[l][k] = b[SPI_BLOB_BASE_LUMPS + 10 * l + k];
  for (int leg = 0; leg < 4; leg++) {
    LegK& L = M->leg[leg];
    for (int j = 0; j < 3; j++) {
      const float* p = b + SPI_BLOB_LEG_BODIES + SPI_LEG_BODY_STRIDE * (3 * leg + j);
      if ((int)p[13] != (j == 0 ? 0 : 1)) return -1;
      const float m = p[0];
      const float c[3] = {p[1], p[2], p[3]};
      const float cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
      L.m[j] = m;
      for (int k = 0; k < 3; k++) L.h[j][k] = m * c[k];
      const float Ic[6] = {p[4], p[5], p[6], p[7], p[8], p[9]};
      for (int r = 0; r < 3; r++)
        for (int k = r; k < 3; k++) {
          const int s = sidx(r, k);
          L.Io[j][s] = Ic[s] + m * ((r == k ? cc : 0.f) - c[r] * c[k]);
        }
      L.qdef[j] = b[SPI_BLOB_Q_DEFAULT + 3 * leg + j];
      L.tlim[j] = b[SPI_BLOB_TORQUE_LIMIT + 3 * leg + j];
      const float* r = p + 10;
      if (j == 0) { if (r[2] != 0.f) return -2; L.rh[0] = r[0]; L.rh[1] = r[1]; }
      if (j == 1) { if (r[0] != 0.f || r[2] != 0.f) return -2; L.rt = r[1]; }
      if (j == 2) { if (r[0] != 0.f || r[1] != 0.f) return -2; L.rc = r[2]; }
    }
    for (int k = 0; k < 3; k++) L.foot[k] = b[SPI_BLOB_FOOT_OFFSET + 3 * leg + k];
    // constant projection of the calf joint (axis y: a = 1, b = 2, c = 0)
    {
      const int a = 1, bb = 2, cx = 0;
      ABI A;
      {
        const float* h = L.h[2];
        for (int i = 0; i < 6; i++) A.I[i] = L.Io[2][i];
        A.H[0] = 0.f; A.H[1] = -h[2]; A.H[2] = h[1]; A.H[3] = h[2]; A.H[4] = 0.f; A.H[5] = -h[0];
        A.H[6] = -h[1]; A.H[7] = h[0]; A.H[8] = 0.f;
        A.M[0] = A.M[1] = A.M[2] = L.m[2]; A.M[3] = A.M[4] = A.M[5] = 0.f;
      }
      for (int i = 0; i < 3; i++) { L.cUa[i] = A.I[sidx(i, a)]; L.cUl[i] = A.H[3 * a + i]; }
      L.cDinv = 1.0f / L.cUa[a];
      float Uad[3], Uld[3];
      for (int i = 0; i < 3; i++) { Uad[i] = L.cUa[i] * L.cDinv; Uld[i] = L.cUl[i] * L.cDinv; }
      L.cUadb = Uad[bb]; L.cUadc = Uad[cx];
      for (int i = 0; i < 3; i++) L.cUld[i] = Uld[i];
      L.cIbb = A.I[sidx(bb, bb)] - Uad[bb] * L.cUa[bb];
      L.cIbc = A.I[sidx(bb, cx)] - Uad[bb] * L.cUa[cx];
      L.cIcc = A.I[sidx(cx, cx)] - Uad[cx] * L.cUa[cx];
      for (int j = 0; j < 3; j++) {
        L.cHb[j] = A.H[3 * bb + j] - Uad[bb] * L.cUl[j];
        L.cHc[j] = A.H[3 * cx + j] - Uad[cx] * L.cUl[j];
      }
      for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++) L.cMa[sidx(i, j)] = A.M[sidx(i, j)] - Uld[i] * L.cUl[j];
    }
  }
  for (int j = 0; j < 12; j++) { M->kp[j] = b[SPI_BLOB_KP + j]; M->kd[j] = b[SPI_BLOB_KD + j]; }
#if defined(SPI_WS_CALF_HARMONIC)
  // harmonic coefficients of the thigh's articulated inertia in the calf angle: discrete Fourier sums (exact for a
  // trigonometric polynomial of degree 2 sampled at N > 4 equispaced angles) of the direct evaluation, accumulated in double
  for (int leg = 0; leg < 4; leg++) {
    LegK& L = M->leg[leg];
    constexpr int N = 32;
    double acc[kCalfEntries][5];
    for (int e = 0; e < kCalfEntries; e++) for (int h = 0; h < 5; h++) acc[e][h] = 0.0;
    for (int k = 0; k < N; k++) {
      const double q = 6.283185307179586 * k / N;
      ABI A2;
      abi_from_rigid(L.m[1], L.h[1], L.Io[1], A2);
      Twist p3, p2;
      for (int i = 0; i < 3; i++) { p3.a[i] = p3.l[i] = 0.f; p2.a[i] = p2.l[i] = 0.f; }
      Keep kk = Keep();
      kk.cs = (float)std::cos(q); kk.sn = (float)std::sin(q);
      calf_inward(L, p3, 0.f, kk, A2, p2);
      const double basis[5] = {1.0, std::cos(q), std::sin(q), std::cos(2 * q), std::sin(2 * q)};
      for (int e = 0; e < kCalfEntries; e++)
        for (int h = 0; h < 5; h++) acc[e][h] += (double)abi_entry(A2, e) * basis[h] * (h == 0 ? 1.0 : 2.0) / N;
    }
    for (int i = 0; i < 64; i++) L.calfA2[i] = 0.f;
    for (int e = 0; e < kCalfEntries; e++) {
      double scale = 1e-3;      // entries of a block share a physical scale: compare against the largest one of the block
      const int e0 = e < 6 ? 0 : (e < 15 ? 6 : 15), e1 = e < 6 ? 6 : (e < 15 ? 15 : 21);
      for (int i = e0; i < e1; i++) for (int h = 0; h < 5; h++) scale = std::fmax(scale, std::fabs(acc[i][h]));
      for (int h = 0; h < 5; h++) {
        if ((calf_mask(e) >> h) & 1) L.calfA2[calf_offset(e, h)] = (float)acc[e][h];
        else if (std::fabs(acc[e][h]) > 2e-6 * scale) return -3;      // not the harmonic structure this path is compiled for
      }
    }
  }
#endif
  return 0;
}

}  // namespace ws
