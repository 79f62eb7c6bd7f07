// fft_regs.cuh -- in-register complex FFT butterflies of size 2,3,4,5,8,16 (fp32).
// Natural-order in, natural-order out; INV=false: exp(-2*pi*i*jk/R) (cv::dft forward,
// SURVEY 8c R1), INV=true: exp(+...), unscaled.  All indices are compile-time so the
// arrays live in registers.
#pragma once
#include <cuda_runtime.h>
#include <type_traits>

namespace fpm {

// Complex arithmetic on float2.  On sm_100a the packed-fp32 pipe is used (FADD2 / FMUL2 / FFMA2 operate on a
// 64-bit register pair; half-swap, per-half negation and scalar broadcast are operand modifiers, so a multiply by
// -i or the (x,x)/(y,y) broadcasts of a complex product cost nothing): half the issue slots of scalar code.
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
#define FPM_PACKED 1
#else
#define FPM_PACKED 0
#endif
__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) {
#if FPM_PACKED
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) {
#if FPM_PACKED
  return __fadd2_rn(a, make_float2(-b.x, -b.y));
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// a * b
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
#if FPM_PACKED
  return __ffma2_rn(make_float2(a.x, a.x), b, __fmul2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x)));
#else
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
#endif
}
// a * conj(b)
__host__ __device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
#if FPM_PACKED
  return __ffma2_rn(make_float2(a.x, a.x), make_float2(b.x, -b.y), __fmul2_rn(make_float2(a.y, a.y), make_float2(b.y, b.x)));
#else
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
#endif
}
// a * s (real scale)
__host__ __device__ __forceinline__ float2 cscale(float2 a, float s) {
#if FPM_PACKED
  return __fmul2_rn(a, make_float2(s, s));
#else
  return make_float2(a.x * s, a.y * s);
#endif
}
// s * a + b (real scalar s): one packed FFMA2 with the scalar broadcast to both halves
__host__ __device__ __forceinline__ float2 cfma(float s, float2 a, float2 b) {
#if FPM_PACKED
  return __ffma2_rn(make_float2(s, s), a, b);
#else
  return make_float2(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y));
#endif
}
// multiply by +i
__host__ __device__ __forceinline__ float2 muli(float2 a) { return make_float2(-a.y, a.x); }
// multiply by -i (forward quarter turn) or +i (inverse)
template <bool INV> __host__ __device__ __forceinline__ float2 rot90(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// multiply by twiddle w (forward table) or its conjugate
template <bool INV> __host__ __device__ __forceinline__ float2 twmul(float2 a, float2 w) {
  return INV ? cmulc(a, w) : cmul(a, w);
}

// a * w from a table entry t = (w.x, s*w.y, -s*w.y, w.x): s = +1 gives a*w, s = -1 gives a*conj(w).  The same two
// packed operations as cmul / cmulc without the half negation they need for a plain (w.x, w.y) operand.
__host__ __device__ __forceinline__ float2 twmul4(float2 a, float4 t) {
#if FPM_PACKED
  return __ffma2_rn(make_float2(a.x, a.x), make_float2(t.x, t.y), __fmul2_rn(make_float2(a.y, a.y), make_float2(t.z, t.w)));
#else
  return make_float2(fmaf(a.x, t.x, a.y * t.z), fmaf(a.x, t.y, a.y * t.w));
#endif
}

template <bool INV> __host__ __device__ __forceinline__ void fft2(float2& a, float2& b) {
  float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

template <bool INV> __host__ __device__ __forceinline__ void fft4(float2& x0, float2& x1, float2& x2, float2& x3) {
  float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = rot90<INV>(csub(x1, x3));
  x0 = cadd(t0, t2);
  x2 = csub(t0, t2);
  x1 = cadd(t1, t3);
  x3 = csub(t1, t3);
}

template <bool INV> __host__ __device__ __forceinline__ void fft3(float2& x0, float2& x1, float2& x2) {
  const float s = INV ? 0.86602540378443864676f : -0.86602540378443864676f;   // Im(W3)
  const float2 t1 = cadd(x1, x2);
  const float2 t2 = cfma(-0.5f, t1, x0);
  const float2 t3 = muli(cscale(csub(x1, x2), s));      // i*s*d
  x0 = cadd(x0, t1);
  x1 = cadd(t2, t3);
  x2 = csub(t2, t3);
}

// packed form (FADD2 / FFMA2 / FMUL2 on whole complex values: 20 instructions instead of ~40 scalar ones; the stages of
// the mixed-radix kernels are issue-bound on exactly these butterflies)
template <bool INV> __host__ __device__ __forceinline__ void fft5(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4) {
  const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;    // cos(2pi/5), cos(4pi/5)
  const float s1 = INV ? 0.95105651629515357212f : -0.95105651629515357212f;  // +-sin(2pi/5)
  const float s2 = INV ? 0.58778525229247312917f : -0.58778525229247312917f;  // +-sin(4pi/5)
  const float2 a1 = cadd(x1, x4), b1 = csub(x1, x4), a2 = cadd(x2, x3), b2 = csub(x2, x3);
  const float2 y0 = cadd(x0, cadd(a1, a2));
  const float2 p1 = cfma(c1, a1, cfma(c2, a2, x0));
  const float2 p2 = cfma(c2, a1, cfma(c1, a2, x0));
  // q = i*(s1*b1 + s2*b2),  r = i*(s2*b1 - s1*b2)
  const float2 q = muli(cfma(s1, b1, cscale(b2, s2)));
  const float2 r = muli(cfma(s2, b1, cscale(b2, -s1)));
  x0 = y0;
  x1 = cadd(p1, q);
  x4 = csub(p1, q);
  x2 = cadd(p2, r);
  x3 = csub(p2, r);
}

// (1 -+ i)/sqrt2 and (-1 -+ i)/sqrt2 multiplications: one add of a with its rotated self, one scale
template <bool INV> __host__ __device__ __forceinline__ float2 mulW8_1(float2 a) {
  const float h = 0.70710678118654752440f;
  // forward: h*((x+y), (y-x)) = h*(a + rot(-i)a) ; inverse: h*((x-y), (x+y)) = h*(a + rot(+i)a)
  return cscale(cadd(a, rot90<INV>(a)), h);
}
template <bool INV> __host__ __device__ __forceinline__ float2 mulW8_3(float2 a) {
  const float h = 0.70710678118654752440f;
  // forward: h*((y-x), -(x+y)) = h*(rot(-i)a - a) ; inverse: h*(-(x+y), (x-y)) = h*(rot(+i)a - a)
  return cscale(csub(rot90<INV>(a), a), h);
}

template <bool INV> __host__ __device__ __forceinline__ void fft8(float2 (&v)[8]) {
  // DIT: even / odd quarter transforms, then W8^k combine
  fft4<INV>(v[0], v[2], v[4], v[6]);
  fft4<INV>(v[1], v[3], v[5], v[7]);
  // after fft4 the outputs sit in slots (0,2,4,6) = E0..E3 and (1,3,5,7) = O0..O3
  float2 o0 = v[1], o1 = mulW8_1<INV>(v[3]), o2 = rot90<INV>(v[5]), o3 = mulW8_3<INV>(v[7]);
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

// multiply by W16^n (forward) or its conjugate, n compile-time in [0,9]
template <bool INV, int n> __host__ __device__ __forceinline__ float2 mulW16(float2 a) {
  if constexpr (n == 0) return a;
  else if constexpr (n == 2) return mulW8_1<INV>(a);
  else if constexpr (n == 4) return rot90<INV>(a);
  else if constexpr (n == 6) return mulW8_3<INV>(a);
  else {
    constexpr float c = (n == 1) ? 0.92387953251128675613f : (n == 3) ? 0.38268343236508977173f : -0.92387953251128675613f;  // n==9
    constexpr float s = (n == 1) ? -0.38268343236508977173f : (n == 3) ? -0.92387953251128675613f : 0.38268343236508977173f;
    return twmul<INV>(a, make_float2(c, s));
  }
}

// second half of the radix-4 x radix-4 DIT 16-point transform: A_j[k] (j = sub-transform, k = its output) sits in
// v[j + 4k]; twiddle by W16^(j*k), then for every k a 4-point transform over j
template <bool INV> __host__ __device__ __forceinline__ void fft16_tail(float2 (&v)[16]) {
  v[5] = mulW16<INV, 1>(v[5]);   v[6] = mulW16<INV, 2>(v[6]);    v[7] = mulW16<INV, 3>(v[7]);
  v[9] = mulW16<INV, 2>(v[9]);   v[10] = mulW16<INV, 4>(v[10]);  v[11] = mulW16<INV, 6>(v[11]);
  v[13] = mulW16<INV, 3>(v[13]); v[14] = mulW16<INV, 6>(v[14]);  v[15] = mulW16<INV, 9>(v[15]);
  // ... then for every k a 4-point transform over j gives X[k + 4q], q = 0..3 (in slots 4k+q)
  fft4<INV>(v[0], v[1], v[2], v[3]);
  fft4<INV>(v[4], v[5], v[6], v[7]);
  fft4<INV>(v[8], v[9], v[10], v[11]);
  fft4<INV>(v[12], v[13], v[14], v[15]);
  // slot 4k+q holds X[k+4q]  ->  transpose the 4x4 index grid to natural order
  float2 t;
#define FPM_SWAP(a, b) t = v[a]; v[a] = v[b]; v[b] = t;
  FPM_SWAP(1, 4) FPM_SWAP(2, 8) FPM_SWAP(3, 12) FPM_SWAP(6, 9) FPM_SWAP(7, 13) FPM_SWAP(11, 14)
#undef FPM_SWAP
}

template <bool INV> __host__ __device__ __forceinline__ void fft16(float2 (&v)[16]) {
  // radix-4 DIT: four stride-4 sub-transforms A_j (j = 0..3) ...
  fft4<INV>(v[0], v[4], v[8], v[12]);
  fft4<INV>(v[1], v[5], v[9], v[13]);
  fft4<INV>(v[2], v[6], v[10], v[14]);
  fft4<INV>(v[3], v[7], v[11], v[15]);
  fft16_tail<INV>(v);
}

// 16-point transform of an input whose entries 3..12 are zero: w = (x[0], x[1], x[2], x[13], x[14], x[15]).
// The first radix-4 layer of fft16 degenerates (8 complex additions instead of 32); the result equals fft16 of the
// zero-padded input bit for bit (x + 0 and 0 - x are exact), up to the sign of exact zeros.
template <bool INV> __host__ __device__ __forceinline__ void fft16_in6(const float2 (&w)[6], float2 (&v)[16]) {
  const float2 x0 = w[0], x1 = w[1], x2 = w[2], x13 = w[3], x14 = w[4], x15 = w[5];
  v[0] = x0; v[4] = x0; v[8] = x0; v[12] = x0;                       // j = 0: (x0, 0, 0, 0)
  {                                                                  // j = 1: (x1, 0, 0, x13)
    const float2 r = rot90<INV>(x13);
    v[1] = cadd(x1, x13); v[5] = csub(x1, r); v[9] = csub(x1, x13); v[13] = cadd(x1, r);
  }
  {                                                                  // j = 2: (x2, 0, 0, x14)
    const float2 r = rot90<INV>(x14);
    v[2] = cadd(x2, x14); v[6] = csub(x2, r); v[10] = csub(x2, x14); v[14] = cadd(x2, r);
  }
  {                                                                  // j = 3: (0, 0, 0, x15)
    const float2 r = rot90<INV>(x15);
    v[3] = x15; v[7] = make_float2(-r.x, -r.y); v[11] = make_float2(-x15.x, -x15.y); v[15] = r;
  }
  fft16_tail<INV>(v);
}

// ---- compile-time loop and 32nd roots of unity -------------------------------------------
template <int B, int E, class F> __host__ __device__ __forceinline__ void static_for(F&& f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    static_for<B + 1, E>(f);
  }
}

__host__ __device__ constexpr float cos32(int n) {
  constexpr float t[32] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654757f,
                           0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f, 0.f, -0.19509032201612819f,
                           -0.38268343236508973f, -0.55557023301960196f, -0.70710678118654746f, -0.83146961230254535f,
                           -0.92387953251128674f, -0.98078528040323043f, -1.f, -0.98078528040323043f, -0.92387953251128685f,
                           -0.83146961230254546f, -0.70710678118654768f, -0.55557023301960218f, -0.38268343236509034f,
                           -0.19509032201612866f, 0.f, 0.1950903220161283f, 0.38268343236509f, 0.55557023301960184f,
                           0.70710678118654735f, 0.83146961230254524f, 0.92387953251128652f, 0.98078528040323032f};
  return t[n & 31];
}
__host__ __device__ constexpr float sin32(int n) { return cos32(n + 24); }   // sin(x) = cos(x - pi/2)

// a * exp(-2*pi*i*n/32)  (INV: exp(+...)), n compile-time
template <bool INV, int n> __host__ __device__ __forceinline__ float2 mulW32(float2 a) {
  constexpr int m = n & 31;
  if constexpr (m == 0) return a;
  else if constexpr (m == 8) return rot90<INV>(a);
  else if constexpr (m == 16) return make_float2(-a.x, -a.y);
  else if constexpr (m == 24) return rot90<!INV>(a);
  else if constexpr (m == 4) return mulW8_1<INV>(a);
  else if constexpr (m == 12) return mulW8_3<INV>(a);
  else {
    constexpr float c = cos32(m), s = -sin32(m);      // forward twiddle c + i*s
    return twmul<INV>(a, make_float2(c, s));
  }
}

template <bool INV> __host__ __device__ __forceinline__ void fft32(float2 (&v)[32]) {
  // radix-4 x radix-8 DIT: four stride-4 sub-transforms A_j (8 points each) ...
  static_for<0, 4>([&](auto J) {
    constexpr int j = decltype(J)::value;
    float2 t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = v[j + 4 * k];
    fft8<INV>(t);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[j + 4 * k] = t[k];   // A_j[k] in slot j + 4k
  });
  // ... twiddled by W32^(j*k) ...
  static_for<1, 4>([&](auto J) {
    constexpr int j = decltype(J)::value;
    static_for<1, 8>([&](auto K) {
      constexpr int k = decltype(K)::value;
      v[j + 4 * k] = mulW32<INV, j * k>(v[j + 4 * k]);
    });
  });
  // ... then eight 4-point transforms over j give X[k + 8q] in slot 4k + q
#pragma unroll
  for (int k = 0; k < 8; ++k) fft4<INV>(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  // slot 4k+q -> natural index k + 8q
  float2 o[32];
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int q = 0; q < 4; ++q) o[k + 8 * q] = v[4 * k + q];
#pragma unroll
  for (int n = 0; n < 32; ++n) v[n] = o[n];
}

template <int R, bool INV> __host__ __device__ __forceinline__ void fftR(float2 (&v)[R]) {
  if constexpr (R == 32) fft32<INV>(v);
  else if constexpr (R == 16) fft16<INV>(v);
  else if constexpr (R == 8) fft8<INV>(v);
  else if constexpr (R == 4) fft4<INV>(v[0], v[1], v[2], v[3]);
  else if constexpr (R == 2) fft2<INV>(v[0], v[1]);
  else static_assert(R == 16, "unsupported radix");
}

// ---- composite radices 6, 9, 10, 12, 15, 20, 25 (mixed-radix tiles: Np = 90 = 9 x 10, 100 = 10 x 10, 200 = 20 x 10 ...) --------------------------
// cos / sin by Taylor series in double, for compile-time roots of unity (|x| <= pi: 17 terms reach 1e-16)
__host__ __device__ constexpr double cx_cos(double x) {
  double s = 1.0, t = 1.0;
  for (int k = 1; k <= 17; ++k) { t *= -x * x / ((2.0 * k - 1.0) * (2.0 * k)); s += t; }
  return s;
}
__host__ __device__ constexpr double cx_sin(double x) {
  double s = x, t = x;
  for (int k = 1; k <= 17; ++k) { t *= -x * x / ((2.0 * k) * (2.0 * k + 1.0)); s += t; }
  return s;
}
// angle of the m-th R-th root of unity, folded into (-pi, pi]
__host__ __device__ constexpr double root_angle(int m, int R) {
  int q = m % R;
  if (2 * q > R) q -= R;
  return 6.283185307179586476925286766559 * q / R;
}

// R = A * B Cooley-Tukey split; B == 1: a direct butterfly
template <int R> struct RadixSplit { static constexpr int A = R, B = 1; };
template <> struct RadixSplit<6> { static constexpr int A = 2, B = 3; };
template <> struct RadixSplit<9> { static constexpr int A = 3, B = 3; };
template <> struct RadixSplit<10> { static constexpr int A = 2, B = 5; };
template <> struct RadixSplit<12> { static constexpr int A = 4, B = 3; };
template <> struct RadixSplit<15> { static constexpr int A = 3, B = 5; };
template <> struct RadixSplit<20> { static constexpr int A = 4, B = 5; };
template <> struct RadixSplit<25> { static constexpr int A = 5, B = 5; };
// fft_reg<R> leaves X[radix_out<R>(i)] in v[i] (slot B*k1 + k2 holds X[k1 + A*k2]); the caller folds the permutation
// into its store addresses
template <int R> __host__ __device__ constexpr int radix_out(int i) {
  return RadixSplit<R>::B == 1 ? i : i / RadixSplit<R>::B + RadixSplit<R>::A * (i % RadixSplit<R>::B);
}

// n-point butterfly over v[O], v[O+S], ...
template <int n, bool INV, int O, int S, int R> __host__ __device__ __forceinline__ void fft_strided(float2 (&v)[R]) {
  if constexpr (n == 2) fft2<INV>(v[O], v[O + S]);
  else if constexpr (n == 3) fft3<INV>(v[O], v[O + S], v[O + 2 * S]);
  else if constexpr (n == 4) fft4<INV>(v[O], v[O + S], v[O + 2 * S], v[O + 3 * S]);
  else if constexpr (n == 5) fft5<INV>(v[O], v[O + S], v[O + 2 * S], v[O + 3 * S], v[O + 4 * S]);
  else static_assert(n == 2, "unsupported butterfly");
}

template <int R, bool INV> __host__ __device__ __forceinline__ void fft_reg(float2 (&v)[R]) {
  constexpr int A = RadixSplit<R>::A, B = RadixSplit<R>::B;
  if constexpr (B == 1) {
    if constexpr (R == 8) fft8<INV>(v);
    else if constexpr (R == 16) fft16<INV>(v);
    else fft_strided<R, INV, 0, 1>(v);
  } else {
    // input index n = B*a + b: for every b an A-point transform over a (k1 lands in slot B*k1 + b) ...
    static_for<0, B>([&](auto Bq) { fft_strided<A, INV, decltype(Bq)::value, B>(v); });
    // ... twiddled by W_R^(b*k1) ...
    static_for<1, A>([&](auto K1) {
      static_for<1, B>([&](auto Bq) {
        constexpr int k1 = decltype(K1)::value, b = decltype(Bq)::value;
        constexpr float c = (float)cx_cos(root_angle(k1 * b, R)), s = (float)cx_sin(root_angle(k1 * b, R));
        v[B * k1 + b] = twmul<INV>(v[B * k1 + b], make_float2(c, -s));
      });
    });
    // ... then for every k1 a B-point transform over b: slot B*k1 + k2 = X[k1 + A*k2]
    static_for<0, A>([&](auto K1) { fft_strided<B, INV, B * decltype(K1)::value, 1>(v); });
  }
}

// ---- pruned butterflies for R = 4 * B (radix 20 of the Np = 200 plan) -------------------------------------------------
// Input pruning: only x[0], x[1], x[2], x[R-3], x[R-2], x[R-1] are non-zero (w = those six, in that order) -- the
// samples of a line that lie inside a narrow pupil box.  The first layer (4-point transforms over a, n = B*a + b)
// sees at most the entries a = 0 and a = 3 and degenerates to copies / one addition; the result equals fft_reg<R> of
// the zero-padded input up to the sign of exact zeros.  Same slot convention as fft_reg<R>.
template <int R, bool INV> __host__ __device__ __forceinline__ void fft_reg_in6(const float2 (&w)[6], float2 (&v)[R]) {
  constexpr int A = RadixSplit<R>::A, B = RadixSplit<R>::B;
  static_assert(A == 4 && B >= 3 && B <= 6, "fft_reg_in6: R = 4 * B, 3 <= B <= 6");
  static_for<0, B>([&](auto Bq) {
    constexpr int b = decltype(Bq)::value;
    constexpr bool has_lo = b <= 2, has_hi = b >= B - 3;
    if constexpr (has_lo && has_hi) {
      const float2 lo = w[b], hi = w[3 + b - (B - 3)], r = rot90<INV>(hi);
      v[b] = cadd(lo, hi); v[B + b] = csub(lo, r); v[2 * B + b] = csub(lo, hi); v[3 * B + b] = cadd(lo, r);
    } else if constexpr (has_lo) {
      const float2 lo = w[b];
      v[b] = lo; v[B + b] = lo; v[2 * B + b] = lo; v[3 * B + b] = lo;
    } else {
      const float2 hi = w[3 + b - (B - 3)], r = rot90<INV>(hi);
      v[b] = hi; v[B + b] = make_float2(-r.x, -r.y); v[2 * B + b] = make_float2(-hi.x, -hi.y); v[3 * B + b] = r;
    }
  });
  static_for<1, A>([&](auto K1) {
    static_for<1, B>([&](auto Bq) {
      constexpr int k1 = decltype(K1)::value, b = decltype(Bq)::value;
      constexpr float c = (float)cx_cos(root_angle(k1 * b, R)), s = (float)cx_sin(root_angle(k1 * b, R));
      v[B * k1 + b] = twmul<INV>(v[B * k1 + b], make_float2(c, -s));
    });
  });
  static_for<0, A>([&](auto K1) { fft_strided<B, INV, B * decltype(K1)::value, 1>(v); });
}

// Output pruning (R = 20): only X[0], X[1], X[2], X[R-3], X[R-2], X[R-1] are wanted (o = those six, in that order) --
// the samples of a transformed line inside a narrow pupil box.  X[k1 + 4*k2]: the last layer (5-point transforms over
// b for every k1) is asked for k2 = 0 (k1 = 0, 1, 2) and k2 = 4 (k1 = 1, 2, 3) only.
template <int R, bool INV> __host__ __device__ __forceinline__ void fft_reg_out6(float2 (&v)[R], float2 (&o)[6]) {
  constexpr int A = RadixSplit<R>::A, B = RadixSplit<R>::B;
  static_assert(A == 4 && B == 5, "fft_reg_out6: R = 20");
  static_for<0, B>([&](auto Bq) { fft_strided<A, INV, decltype(Bq)::value, B>(v); });
  static_for<1, A>([&](auto K1) {
    static_for<1, B>([&](auto Bq) {
      constexpr int k1 = decltype(K1)::value, b = decltype(Bq)::value;
      constexpr float c = (float)cx_cos(root_angle(k1 * b, R)), s = (float)cx_sin(root_angle(k1 * b, R));
      v[B * k1 + b] = twmul<INV>(v[B * k1 + b], make_float2(c, -s));
    });
  });
  const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;    // cos(2pi/5), cos(4pi/5)
  const float s1 = INV ? 0.95105651629515357212f : -0.95105651629515357212f;  // +-sin(2pi/5)
  const float s2 = INV ? 0.58778525229247312917f : -0.58778525229247312917f;  // +-sin(4pi/5)
  static_for<0, A>([&](auto K1) {
    constexpr int k1 = decltype(K1)::value;
    const float2 y0 = v[B * k1], y1 = v[B * k1 + 1], y2 = v[B * k1 + 2], y3 = v[B * k1 + 3], y4 = v[B * k1 + 4];
    const float2 a1 = cadd(y1, y4), a2 = cadd(y2, y3);
    if constexpr (k1 <= 2) o[k1] = cadd(y0, cadd(a1, a2));                                   // k2 = 0
    if constexpr (k1 >= 1) {                                                                 // k2 = 4: p1 - q of fft5
      const float2 b1 = csub(y1, y4), b2 = csub(y2, y3);
      const float2 p1 = cfma(c1, a1, cfma(c2, a2, y0));
      const float2 q = muli(cfma(s1, b1, cscale(b2, s2)));
      o[2 + k1] = csub(p1, q);
    }
  });
}

}  // namespace fpm
