// rtb_pretest.h -- conservative ray/triangle REJECTION test that precedes the exact one in the throughput kernels.
//
// The exact test (rtb_device.cuh: triIntersectT, reference Triangle.cpp:38-121) is Cramer's rule on four 3x3
// determinants, each evaluated in the reference's association without FMA, followed by three IEEE divisions:
// ~105-140 SASS instructions per test, 65 % of the instructions of a tunnel frame.  Almost every test ends in a
// reject.  sureReject() below decides most of those rejects from the SAME four determinants evaluated the cheap
// way (two cross products and four dot products, FMA, no division: 27 flops) plus a running bound on how far the
// cheap and the reference values can be apart.  It never accepts anything: whatever it cannot reject with
// certainty is a "candidate" and goes through the exact test unchanged, so hit ids, distances and images stay
// bit-identical to the reference.  This header is plain C++ (host + device): tests/ evaluates the same function
// on the CPU over every (ray, triangle) pair of whole frames and checks it against the oracle's exact test.
//
// Notation.  Float inputs, identical on both sides: e1 = a - b, e2 = a - c (Triangle.cpp:73-79), d = ray direction,
// b = a - origin (Triangle.cpp:81-83).  Exact real determinants of those floats:
//      DetM = det[e1 e2 d]   DetT = det[e1 e2 b]   DetB = det[b e2 d]   DetG = det[e1 b d]
// and beta = DetB / DetM, gamma = DetG / DetM, t = DetT / DetM.  With u = 2^-24 and S_X the sum of the absolute
// values of the six triple products of Det X:
//   * the reference evaluates every Det X as a sum of six float triple products (2 roundings per product, 5 for the
//     sum): |det_ref - Det| <= gamma_7 S_X                                   (gamma_k = k u / (1 - k u))
//   * here P = e2 x d, Q = b x e1 (one multiply + one FMA per component: 2 roundings per product), then
//     detM' = e1.P, detB' = b.P, detT' = e2.Q, detG' = -(d.Q) (three more roundings): |det' - Det| <= gamma_5 S_X
//   so |det' - det_ref| <= 12.1 u S_X < 2^-20 S_X; RTB_PRE_EPS = 2^-19 leaves a factor 2.6 for the roundings of the
//   bounds themselves (each a handful of float operations, relative error < 1e-6).  Measured against exact (binary128)
//   determinants on 10 M random pairs: the two gamma claims hold with 0.70 / 0.72 of their limit used, and the largest
//   |det' - det_ref| is 0.07 of the kappa the decisions use (tests/test_pretest.py, profiles/r01_pretest_stats.md).
// Cheap upper bounds of S_X, with dmx = max |d_i|, bn = |b_x| + |b_y| + |b_z| and two per-triangle constants
// A1 = sum over i != j of |e1_i| |e2_j|  and  E = max(|e1|_1, |e2|_1), both rounded up on upload:
//      S_M <= dmx A1      S_B, S_G <= dmx bn E      S_T <= bn A1
// hence  kappa = EPS dmx (A1 + bn E) >= |detX' - detX_ref| for X = M, B, G   and   kT = EPS bn A1 for X = T.
//
// Decisions (m = |detM'|, s = sign(detM'), X~ = s detX'): only if m > kappa -- then the reference's det_M has the
// sign s and |det_M| in [m - kappa, m + kappa] -- any of
//      B~ < -(CL (m + kappa) + kappa)         =>  beta_ref < -1e-4          (Triangle.cpp:102)
//      B~ >   CH (m + kappa) + kappa          =>  beta_ref > 1.0001
//      the same two for G~                    =>  gamma_ref outside its window (Triangle.cpp:108)
//      B~ + G~ > CH (m + kappa) + 2 kappa     =>  beta_ref, gamma_ref in their windows imply 1 - beta - gamma < -1e-4
//      T~ + kT < L' (m - kappa)               =>  t_ref < L   (L = max(0.0005, leaf window low) > 0, L' = L (1 - 1e-6))
//      T~ - kT > H' (m + kappa)               =>  t_ref > H   (H = max(0, min(leaf window high, nearest hit so far)),
//                                                              H' = H (1 + 1e-6))
// proves that the reference rejects (the float quotient is a monotonic rounding of the real quotient; CL, CH carry
// a margin of 1e-8 resp. 1e-5 relative, far above u).  A NaN anywhere makes every comparison false: candidate.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RTB_PRE_FN __host__ __device__ __forceinline__
#else
#define RTB_PRE_FN inline
#endif

namespace rtb_pre {

#define RTB_PRE_EPS 1.9073486328125e-06f // 2^-19
#define RTB_PRE_CL 1.0001e-4f
#define RTB_PRE_CH 1.00011f
#define RTB_PRE_LSCALE 0.999999f
#define RTB_PRE_HSCALE 1.000001f

// 48 bytes, three 128-bit loads: {a.xyz, A1 EPS} {e1.xyz, E EPS} {e2.xyz, 0}
struct PreTri
{
    float ax, ay, az, a1e;
    float e1x, e1y, e1z, ee;
    float e2x, e2y, e2z, pad;
};

// a, b, c as the 9 leading floats of the 12-float triangle record of include/rtb.h.  Host + device: the upload
// packs on the GPU (k_pack_triangles), the CPU checker of tests/ on the host -- IEEE double arithmetic without
// contraction on both sides (-fmad=false / -ffp-contract=off), hence the same constants
RTB_PRE_FN PreTri makePreTri(const float *t)
{
    PreTri p;
    p.ax = t[0]; p.ay = t[1]; p.az = t[2];
    p.e1x = t[0] - t[3]; p.e1y = t[1] - t[4]; p.e1z = t[2] - t[5];
    p.e2x = t[0] - t[6]; p.e2y = t[1] - t[7]; p.e2z = t[2] - t[8];
    const double x1 = fabs((double)p.e1x), y1 = fabs((double)p.e1y), z1 = fabs((double)p.e1z);
    const double x2 = fabs((double)p.e2x), y2 = fabs((double)p.e2y), z2 = fabs((double)p.e2z);
    const double a1 = x1 * (y2 + z2) + y1 * (x2 + z2) + z1 * (x2 + y2);
    const double e = fmax(x1 + y1 + z1, x2 + y2 + z2);
    // rounded up (1 + 1e-6 covers the double -> float conversion); the tiny floor keeps the bound meaningful
    // when products underflow
    p.a1e = (float)(a1 * (double)RTB_PRE_EPS * 1.000001) + 1e-37f;
    p.ee = (float)(e * (double)RTB_PRE_EPS * 1.000001) + 1e-37f;
    p.pad = 0.f;
    return p;
}

RTB_PRE_FN float xorSign(float v, uint32_t sign)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__float_as_uint(v) ^ sign);
#else
    uint32_t b;
    memcpy(&b, &v, 4);
    b ^= sign;
    memcpy(&v, &b, 4);
    return v;
#endif
}
RTB_PRE_FN uint32_t signOf(float v)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(v) & 0x80000000u;
#else
    uint32_t b;
    memcpy(&b, &v, 4);
    return b & 0x80000000u;
#endif
}

// per-ray constant
RTB_PRE_FN float dirMax(float dx, float dy, float dz) { return fmaxf(fabsf(dx), fmaxf(fabsf(dy), fabsf(dz))); }
// per-leaf constants: lo / hi = the k-d leaf window (Tunnel.cpp:1273-1274) or -FLT_MAX / FLT_MAX for a grid cell;
// nearest = nearest accepted hit of the list so far
RTB_PRE_FN float lowBound(float lo) { return fmaxf(0.0005f, lo) * RTB_PRE_LSCALE; }
RTB_PRE_FN float highBound(float hi, float nearest) { return fmaxf(fminf(hi, nearest), 0.f) * RTB_PRE_HSCALE; }

// true: the reference's Triangle::intersect rejects this pair, or accepts it with t < L or t > H.
// false: unknown -- run the exact test.  HIGH = false drops the t > H test (grid cells: no window, H = +inf).
template <bool HIGH = true>
RTB_PRE_FN bool sureReject(const PreTri &T, float ox, float oy, float oz, float dx, float dy, float dz, float dmx,
                           float Lp, float Hp)
{
    const float b1 = T.ax - ox, b2 = T.ay - oy, b3 = T.az - oz;
    const float px = fmaf(T.e2y, dz, -(T.e2z * dy));
    const float py = fmaf(T.e2z, dx, -(T.e2x * dz));
    const float pz = fmaf(T.e2x, dy, -(T.e2y * dx));
    const float dM = fmaf(T.e1z, pz, fmaf(T.e1y, py, T.e1x * px));
    const float dB = fmaf(b3, pz, fmaf(b2, py, b1 * px));
    const float qx = fmaf(b2, T.e1z, -(b3 * T.e1y));
    const float qy = fmaf(b3, T.e1x, -(b1 * T.e1z));
    const float qz = fmaf(b1, T.e1y, -(b2 * T.e1x));
    const float dT = fmaf(T.e2z, qz, fmaf(T.e2y, qy, T.e2x * qx));
    const float dGn = fmaf(dz, qz, fmaf(dy, qy, dx * qx)); // = -detG'
    const float bn = fabsf(b1) + fabsf(b2) + fabsf(b3);
    const float kap = dmx * fmaf(T.ee, bn, T.a1e);
    const float kT = T.a1e * bn;
    const float m = fabsf(dM);
    const uint32_t s = signOf(dM);
    const float Bt = xorSign(dB, s), Gt = xorSign(dGn, s ^ 0x80000000u), Tt = xorSign(dT, s);
    const float mk = m + kap, mlo = m - kap;
    const float hiK = fmaf(RTB_PRE_CH, mk, kap), loK = -fmaf(RTB_PRE_CL, mk, kap);
    bool rej = (Bt < loK) | (Bt > hiK) | (Gt < loK) | (Gt > hiK) | (Bt + Gt > hiK + kap) | (Tt + kT < Lp * mlo);
    if (HIGH) rej = rej | (Tt - kT > Hp * mk);
    return rej & (mlo > 0.f);
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------------------------
// TWO triangles per call on Blackwell's packed FP32 pipe (sm_100: FFMA2 / FMUL2 / FADD2, `fma.rn.f32x2`).
// The list scans are bound by warp-instruction ISSUE, not by the FMA pipe (36 % busy): a packed instruction does two
// lanes' worth of IEEE arithmetic for one issue slot, the ray's components enter as broadcast scalar operands, and
// each half is rounded exactly like the scalar operation -- the two halves of sureReject2 take, bit for bit, the
// decisions sureReject<HIGH> takes (tests/test_gpu_parity.py::test_packed_pretest_equals_scalar), so everything the
// CPU checker proves about the scalar function holds here.
// Operands come from the PAIR stream: for list positions 2p and 2p + 1 of the accelerator's reference array, six
// float4 = {ax ax' ay ay'} {az az' A1e A1e'} {e1x e1x' e1y e1y'} {e1z e1z' Ee Ee'} {e2x e2x' e2y e2y'} {e2z e2z' 0 0}
// (k_pack_pairs): 64-bit aligned register pairs straight out of the 128-bit loads, and no index indirection.
// ---------------------------------------------------------------------------------------------------------------
struct PreTri2 { float2 ax, ay, az, a1e, e1x, e1y, e1z, ee, e2x, e2y, e2z; };

__device__ __forceinline__ float2 pre2(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ float2 bc(float v) { return make_float2(v, v); }             // becomes a broadcast operand
__device__ __forceinline__ float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }      // folded into operand modifiers
__device__ __forceinline__ float2 abs2(float2 v) { return make_float2(fabsf(v.x), fabsf(v.y)); }

template <bool HIGH = true>
__device__ __forceinline__ void sureReject2(const PreTri2 &T, float ox, float oy, float oz, float dx, float dy, float dz, float dmx,
                                            float Lp, float Hp, bool &rej0, bool &rej1)
{
    const float2 b1 = __fadd2_rn(T.ax, bc(-ox)), b2 = __fadd2_rn(T.ay, bc(-oy)), b3 = __fadd2_rn(T.az, bc(-oz));
    const float2 px = __ffma2_rn(T.e2y, bc(dz), neg2(__fmul2_rn(T.e2z, bc(dy))));
    const float2 py = __ffma2_rn(T.e2z, bc(dx), neg2(__fmul2_rn(T.e2x, bc(dz))));
    const float2 pz = __ffma2_rn(T.e2x, bc(dy), neg2(__fmul2_rn(T.e2y, bc(dx))));
    const float2 dM = __ffma2_rn(T.e1z, pz, __ffma2_rn(T.e1y, py, __fmul2_rn(T.e1x, px)));
    const float2 dB = __ffma2_rn(b3, pz, __ffma2_rn(b2, py, __fmul2_rn(b1, px)));
    const float2 qx = __ffma2_rn(b2, T.e1z, neg2(__fmul2_rn(b3, T.e1y)));
    const float2 qy = __ffma2_rn(b3, T.e1x, neg2(__fmul2_rn(b1, T.e1z)));
    const float2 qz = __ffma2_rn(b1, T.e1y, neg2(__fmul2_rn(b2, T.e1x)));
    const float2 dT = __ffma2_rn(T.e2z, qz, __ffma2_rn(T.e2y, qy, __fmul2_rn(T.e2x, qx)));
    const float2 dGn = __ffma2_rn(bc(dz), qz, __ffma2_rn(bc(dy), qy, __fmul2_rn(bc(dx), qx))); // = -detG'
    const float2 bn = __fadd2_rn(__fadd2_rn(abs2(b1), abs2(b2)), abs2(b3));
    const float2 kap = __fmul2_rn(bc(dmx), __ffma2_rn(T.ee, bn, T.a1e));
    const float2 kT = __fmul2_rn(T.a1e, bn);
    const float2 m = abs2(dM);
    const float2 mk = __fadd2_rn(m, kap), mlo = __fadd2_rn(m, neg2(kap));
    const float2 hiK = __ffma2_rn(bc(RTB_PRE_CH), mk, kap), cLo = __ffma2_rn(bc(RTB_PRE_CL), mk, kap); // cLo = -loK
    const float2 hiA = __fadd2_rn(hiK, kap);
    const float2 tLow = __fmul2_rn(bc(Lp), mlo), tHigh = __fmul2_rn(bc(Hp), mk);
    // sign of det_M' transferred by an exact multiplication with +-1 (one packed instruction per determinant instead of
    // one LOP3 per half; same values as xorSign up to the sign of a NaN, which no comparison sees)
    const float2 sg = make_float2(__uint_as_float(signOf(dM.x) | 0x3f800000u), __uint_as_float(signOf(dM.y) | 0x3f800000u));
    const float2 Bt = __fmul2_rn(dB, sg), Gn = __fmul2_rn(dGn, sg) /* = -G~ */, Tt = __fmul2_rn(dT, sg);
    const float2 sBG = __fadd2_rn(Bt, neg2(Gn)), tPlus = __fadd2_rn(Tt, kT), tMinus = __fadd2_rn(Tt, neg2(kT));
#define RTB_PRE_HALF(h, out)                                                                                                  \
    {                                                                                                                         \
        bool r = (Bt.h < -cLo.h) | (Bt.h > hiK.h) | (Gn.h > cLo.h) | (Gn.h < -hiK.h) | (sBG.h > hiA.h) | (tPlus.h < tLow.h);   \
        if (HIGH) r = r | (tMinus.h > tHigh.h);                                                                               \
        out = r & (mlo.h > 0.f);                                                                                              \
    }
    RTB_PRE_HALF(x, rej0)
    RTB_PRE_HALF(y, rej1)
#undef RTB_PRE_HALF
}
#endif // __CUDACC__

} // namespace rtb_pre
