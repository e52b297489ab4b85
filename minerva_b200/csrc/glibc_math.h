/* glibc_math.h -- glibc 2.39's single-precision expf / logf / tanhf (and the reference's sigmoid built on expf), restated
 * operation for operation so that the "exact mode" kernels of elementwise.cu (mnv_*_exact) return the CPU reference's
 * bits.  One text for host and device: the includer defines the arithmetic macros (MNV_MUL ... MNV_FMA: IEEE operations
 * that the compiler must not contract or reassociate; on the device the __d*_rn / __f*_rn / __fma_rn intrinsics).
 *
 * Why: the reference CPU ops call libm (minerva/op/impl/basic.cpp:134 expf, :139 logf, :444 tanhf, :416 sigmoid =
 * (float)(1.0 / (1.0 + (double)expf(-x)))), and north_star asks for bit-exact elementwise / activation outputs.  libm is
 * a third-party dependency that is not in /root/reference: glibc 2.39 (Ubuntu 2.39-0ubuntu8.5 in this image), whose
 * algorithms are
 *   expf, logf : the table + polynomial kernels in double precision from ARM's optimized-routines
 *                (sysdeps/ieee754/flt-32/e_expf.c, e_logf.c; tables e_exp2f_data.c, e_logf_data.c).  On x86-64 with
 *                FMA/AVX2 the ifunc picks the -mfma build (sysdeps/x86_64/fpu/multiarch/e_expf.c): GCC contracts every
 *                a*b+c of these kernels into an fma (including r = InvLn2N*x - kd) -- PINNED by comparing this text, compiled
 *                for the host, with the libm of this image on ALL 2^32 inputs: 0 mismatches for expf, logf, tanhf and the
 *                sigmoid formula (tests/cpp/check_glibc_math.c; without the fused r, expf differs on 2 inputs);
 *   tanhf      : the fdlibm routine (sysdeps/ieee754/flt-32/s_tanhf.c) over expm1f (s_expm1f.c), plain float arithmetic.
 * The tables were re-derived (exp2f: T[i] = bits(2^(i/32)) - (i << 47), checked against the bytes of libm.so.6) or are
 * the published constants of e_logf_data.c (16 (1/c, log c) pairs).
 *
 */
#ifndef MNV_GLIBC_MATH_H_
#define MNV_GLIBC_MATH_H_
#include <stdint.h>

#ifndef MNV_GM_FN
#define MNV_GM_FN static inline
#endif
#ifndef MNV_GM_CONST
#define MNV_GM_CONST static const
#endif

MNV_GM_CONST uint64_t mnv_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull,
    0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull,
    0x3feedea64c123422ull, 0x3feece086061892dull, 0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull,
    0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, 0x3feee89f995ad3adull,
    0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, 0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full,
    0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
/* (1/c, log c) for the 16 sub-intervals of [0x3f330000, 2 * 0x3f330000) */
MNV_GM_CONST double mnv_logf_tab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};

MNV_GM_FN uint32_t mnv_f2u(float f) { union { float f; uint32_t u; } v; v.f = f; return v.u; }
MNV_GM_FN float mnv_u2f(uint32_t u) { union { float f; uint32_t u; } v; v.u = u; return v.f; }
MNV_GM_FN uint64_t mnv_d2u(double d) { union { double d; uint64_t u; } v; v.d = d; return v.u; }
MNV_GM_FN double mnv_u2d(uint64_t u) { union { double d; uint64_t u; } v; v.u = u; return v.d; }

/* e_expf.c: exp(x) = 2^(k/32) * 2^(r/32), k = round(32 x / ln 2) */
MNV_GM_FN float mnv_glibc_expf(float x) {
  const double InvLn2N = 0x1.71547652b82fep+5, Shift = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-20, C1 = 0x1.ebfce50fac4f3p-13, C2 = 0x1.62e42ff0c52d6p-6;   /* poly_scaled */
  const uint32_t abstop = (mnv_f2u(x) >> 20) & 0x7ff;
  if (abstop >= 0x42b) {                       /* |x| >= 88 or x is nan (top12(88.0f) = 0x42b) */
    if (mnv_f2u(x) == 0xff800000u) return 0.0f;
    if (abstop >= 0x7f8) return x + x;
    if (x > 0x1.62e42ep6f) return MNV_INFF;    /* overflow */
    if (x < -0x1.9fe368p6f) return 0.0f;       /* underflow */
  }
  const double xd = (double)x;
  double kd = MNV_FMA(InvLn2N, xd, Shift);      /* both uses of the product are additions: GCC fuses both */
  const uint64_t ki = mnv_d2u(kd);
  kd = MNV_SUB(kd, Shift);
  const double r = MNV_FMA(InvLn2N, xd, -kd);
  uint64_t t = mnv_exp2f_tab[ki % 32];
  t += ki << (52 - 5);
  const double s = mnv_u2d(t);
  const double zz = MNV_FMA(C0, r, C1);
  const double r2 = MNV_MUL(r, r);
  double y = MNV_FMA(C2, r, 1.0);
  y = MNV_FMA(zz, r2, y);
  y = MNV_MUL(y, s);
  return (float)y;
}

/* e_logf.c: log(x) = log1p(z/c - 1) + log(c) + k ln 2 */
MNV_GM_FN float mnv_glibc_logf(float x) {
  const double Ln2 = 0x1.62e42fefa39efp-1, A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
  uint32_t ix = mnv_f2u(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {      /* x < 0x1p-126 or inf or nan */
    if (ix * 2 == 0) return -MNV_INFF;                      /* log(+-0) = -inf */
    if (ix == 0x7f800000u) return x;
    if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return MNV_NANF;
    ix = mnv_f2u(x * 0x1p23f);                              /* subnormal: normalise */
    ix -= 23u << 23;
  }
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> (23 - 4)) % 16);
  const int k = (int32_t)tmp >> 23;
  const uint32_t iz = ix - (tmp & (0x1ffu << 23));
  const double invc = mnv_logf_tab[i][0], logc = mnv_logf_tab[i][1];
  const double z = (double)mnv_u2f(iz);
  const double r = MNV_FMA(z, invc, -1.0);
  const double y0 = MNV_FMA((double)k, Ln2, logc);
  const double r2 = MNV_MUL(r, r);
  double y = MNV_FMA(A1, r, A2);
  y = MNV_FMA(A0, r2, y);
  y = MNV_FMA(y, r2, MNV_ADD(y0, r));
  return (float)y;
}

/* s_expm1f.c (fdlibm), float arithmetic, no contraction */
MNV_GM_FN float mnv_glibc_expm1f(float x) {
  const float huge = 1.0e+30f, tiny = 1.0e-30f, one = 1.0f, o_threshold = 8.8721679688e+01f, ln2_hi = 6.9313812256e-01f,
              ln2_lo = 9.0580006145e-06f, invln2 = 1.4426950216e+00f, Q1 = -3.3333335072e-02f, Q2 = 1.5873016091e-03f,
              Q3 = -7.9365076090e-05f, Q4 = 4.0082177293e-06f, Q5 = -2.0109921195e-07f;
  float y, hi, lo, c = 0.f, t, e, hxs, hfx, r1;
  int32_t k;
  uint32_t hx = mnv_f2u(x);
  const uint32_t xsb = hx & 0x80000000u;
  hx &= 0x7fffffffu;
  if (hx >= 0x4195b844u) {                     /* |x| >= 27 ln 2 */
    if (hx >= 0x42b17218u) {                   /* |x| >= 88.721... */
      if (hx > 0x7f800000u) return x + x;
      if (hx == 0x7f800000u) return xsb == 0 ? x : -1.0f;
      if (x > o_threshold) return MNV_FMULF(huge, huge);
    }
    if (xsb != 0) return MNV_FSUBF(tiny, one);
  }
  if (hx > 0x3eb17218u) {                      /* |x| > 0.5 ln 2 */
    if (hx < 0x3F851592u) {                    /* and |x| < 1.5 ln 2 */
      if (xsb == 0) { hi = MNV_FSUBF(x, ln2_hi); lo = ln2_lo; k = 1; }
      else { hi = MNV_FADDF(x, ln2_hi); lo = -ln2_lo; k = -1; }
    } else {
      k = (int32_t)MNV_FADDF(MNV_FMULF(invln2, x), xsb == 0 ? 0.5f : -0.5f);
      t = (float)k;
      hi = MNV_FSUBF(x, MNV_FMULF(t, ln2_hi));
      lo = MNV_FMULF(t, ln2_lo);
    }
    x = MNV_FSUBF(hi, lo);
    c = MNV_FSUBF(MNV_FSUBF(hi, x), lo);
  } else if (hx < 0x33000000u) {               /* |x| < 2^-25 */
    t = MNV_FADDF(huge, x);
    return MNV_FSUBF(x, MNV_FSUBF(t, MNV_FADDF(huge, x)));
  } else {
    k = 0;
  }
  hfx = MNV_FMULF(0.5f, x);
  hxs = MNV_FMULF(x, hfx);
  r1 = MNV_FADDF(one, MNV_FMULF(hxs, MNV_FADDF(Q1, MNV_FMULF(hxs, MNV_FADDF(Q2, MNV_FMULF(hxs, MNV_FADDF(Q3, MNV_FMULF(hxs, MNV_FADDF(Q4, MNV_FMULF(hxs, Q5)))))))))); 
  t = MNV_FSUBF(3.0f, MNV_FMULF(r1, hfx));
  e = MNV_FMULF(hxs, MNV_FDIVF(MNV_FSUBF(r1, t), MNV_FSUBF(6.0f, MNV_FMULF(x, t))));
  if (k == 0) return MNV_FSUBF(x, MNV_FSUBF(MNV_FMULF(x, e), hxs));
  e = MNV_FSUBF(MNV_FMULF(x, MNV_FSUBF(e, c)), c);
  e = MNV_FSUBF(e, hxs);
  if (k == -1) return MNV_FSUBF(MNV_FMULF(0.5f, MNV_FSUBF(x, e)), 0.5f);
  if (k == 1) {
    if (x < -0.25f) return MNV_FMULF(-2.0f, MNV_FSUBF(e, MNV_FADDF(x, 0.5f)));
    return MNV_FADDF(one, MNV_FMULF(2.0f, MNV_FSUBF(x, e)));
  }
  if (k <= -2 || k > 56) {
    y = MNV_FSUBF(one, MNV_FSUBF(e, x));
    y = mnv_u2f(mnv_f2u(y) + ((uint32_t)k << 23));
    return MNV_FSUBF(y, one);
  }
  if (k < 23) {
    t = mnv_u2f(0x3f800000u - (0x1000000u >> k));         /* 1 - 2^-k */
    y = MNV_FSUBF(t, MNV_FSUBF(e, x));
    y = mnv_u2f(mnv_f2u(y) + ((uint32_t)k << 23));
  } else {
    t = mnv_u2f((uint32_t)(0x7f - k) << 23);              /* 2^-k */
    y = MNV_FSUBF(x, MNV_FADDF(e, t));
    y = MNV_FADDF(y, one);
    y = mnv_u2f(mnv_f2u(y) + ((uint32_t)k << 23));
  }
  return y;
}

/* s_tanhf.c (fdlibm) */
MNV_GM_FN float mnv_glibc_tanhf(float x) {
  const float one = 1.0f, two = 2.0f, tiny = 1.0e-30f;
  float t, z;
  const uint32_t jx = mnv_f2u(x), ix = jx & 0x7fffffffu;
  if (ix >= 0x7f800000u) {
    if ((int32_t)jx >= 0) return MNV_FADDF(MNV_FDIVF(one, x), one);
    return MNV_FSUBF(MNV_FDIVF(one, x), one);
  }
  if (ix < 0x41b00000u) {                      /* |x| < 22 */
    if (ix == 0) return x;
    if (ix < 0x24000000u) return MNV_FMULF(x, MNV_FADDF(one, x));
    const float ax = mnv_u2f(ix);
    if (ix >= 0x3f800000u) {
      t = mnv_glibc_expm1f(MNV_FMULF(two, ax));
      z = MNV_FSUBF(one, MNV_FDIVF(two, MNV_FADDF(t, two)));
    } else {
      t = mnv_glibc_expm1f(MNV_FMULF(-two, ax));
      z = MNV_FDIVF(-t, MNV_FADDF(t, two));
    }
  } else {
    z = MNV_FSUBF(one, tiny);
  }
  return (int32_t)jx >= 0 ? z : -z;
}

/* basic.cpp:416: 1.0 / (1.0 + expf(-x)) evaluated in double, stored to float */
MNV_GM_FN float mnv_ref_sigmoidf(float x) {
  return (float)MNV_DIV(1.0, MNV_ADD(1.0, (double)mnv_glibc_expf(-x)));
}
#endif /* MNV_GLIBC_MATH_H_ */
