// Goldilocks field (p = 2^64 - 2^32 + 1) and its cubic extension F_p[x]/(x^3 - x - 1) for sm_100a.
//
// Semantics follow the reference's F3g (src/helpers/f3g.js:17-104): canonical little-endian u64 words in
// and out.  Internally values may be any 64-bit representative ("lazy" form) -- every entry point that
// writes to memory canonicalises with gl_canon() -- because the integer pipes only need one conditional
// correction per operation that way.  All arithmetic is 32-bit IMAD / IADD3 on the integer pipes; there is
// no floating point and no tensor-core use (these are not dense contractions).
#pragma once
#include <stdint.h>

typedef uint64_t u64;
typedef unsigned int u32;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL          // 2^64 mod p
#define GL_SHIFT 7ULL                 // coset generator (f3g.js:22)
#define GL_W32 7277203076849721926ULL // order-2^32 root of unity (f3g.js:40)

#define GL_HD __host__ __device__ __forceinline__
#define GL_D __device__ __forceinline__

// ---- lazy-form primitives: inputs/outputs are arbitrary u64 representatives ---------------------------
GL_D u64 gl_canon(u64 a) { return a >= GL_P ? a - GL_P : a; }

GL_D u64 gl_add(u64 a, u64 b) {
    // a + b may wrap 2^64 once; 2^64 = EPS (mod p).  After the first correction the sum is < 2^64 unless
    // both inputs were >= p, so a second (rare) wrap is handled too.
    u64 r = a + b;
    if (r < a) {
        r += GL_EPS;
        if (r < GL_EPS) r += GL_EPS;
    }
    return r;
}

GL_D u64 gl_sub(u64 a, u64 b) {
    u64 r = a - b;
    if (a < b) {
        // wrapped value = a - b + 2^64; subtract EPS to get a - b + p (mod 2^64)
        u64 t = r - GL_EPS;
        if (r < GL_EPS) t -= GL_EPS;
        r = t;
    }
    return r;
}

GL_D u64 gl_neg(u64 a) { return gl_sub(0, a); }

// 128 -> 64 bits with 2^64 = 2^32 - 1 and 2^96 = -1 (mod p):
//     hi*2^64 + lo  ==  V = lo + h0*2^32 - h0 - h1     (h0 = low word of hi, h1 = high word of hi)
// V is computed exactly in three 32-bit words (top word v2 in {-1,0,1}) with carry chains, then v2*2^64 = v2*EPS is
// folded once; no second wrap is possible (V <= 2^65 - 2^33 and V >= -(2^32 - 1)).  12 ALU-pipe instructions, no
// multiply and no compare/select pairs (the compiler's version of the same identity costs ~17 incl. one IMAD.WIDE).
GL_D u64 gl_reduce128(u64 hi, u64 lo) {
    const u32 l0 = (u32)lo, l1 = (u32)(lo >> 32), h0 = (u32)hi, h1 = (u32)(hi >> 32);
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 v0, v1, v2, t, m;\n\t"
        "sub.cc.u32  v0, %2, %4;\n\t"
        "subc.cc.u32 v1, %3, 0;\n\t"
        "subc.u32    v2, 0, 0;\n\t"
        "sub.cc.u32  v0, v0, %5;\n\t"
        "subc.cc.u32 v1, v1, 0;\n\t"
        "subc.u32    v2, v2, 0;\n\t"
        "add.cc.u32  v1, v1, %4;\n\t"
        "addc.u32    v2, v2, 0;\n\t"
        "neg.s32     t, v2;\n\t"
        "shr.s32     m, v2, 31;\n\t"
        "add.cc.u32  %0, v0, t;\n\t"
        "addc.u32    %1, v1, m;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(l0), "r"(l1), "r"(h0), "r"(h1));
    return ((u64)r1 << 32) | r0;
}

GL_D u64 gl_mul(u64 a, u64 b) { return gl_reduce128(__umul64hi(a, b), a * b); }
GL_D u64 gl_sqr(u64 a) { return gl_mul(a, a); }

// value < 2^96 given as (hi32 : lo64): hi32*2^64 + lo = lo + hi32*EPS
GL_D u64 gl_reduce96(u32 hi, u64 lo) {
    u64 t1 = (u64)hi * (u64)0xFFFFFFFFu;
    u64 r = lo + t1;
    if (r < t1) r += GL_EPS;
    return r;
}

// ---- Montgomery multiply (R = 2^64) and one-canonical-operand add/sub ------------------------------------------
// sm_100a executes integer work on two half-rate pipes per SM sub-partition: "alu" (IADD3/LOP3/SHF) and "fmaheavy"
// (IMAD, IMAD.WIDE = 2 slots).  A Goldilocks product is 4 IMAD.WIDE; what decides throughput is how many ALU-pipe
// instructions the reduction needs.  The Montgomery reduction below needs 9 (the 2^64 = 2^32 - 1 fold above needs 12) and
// returns a canonical value whenever one factor is canonical, which lets the add/sub that follow skip the second wrap
// check.  Constants (twiddles, round constants, 2^128 mod p) are stored pre-multiplied by 2^64, so that
// gl_mmul(x, c * 2^64) = x * c with no domain change for x.  Measured (tools/probe/gl_probe.cu, B200): butterfly
// 0.54 -> 0.90 T/s, multiply 1.28 -> 1.41 T/s.
//
// NOTE on carry chains: ptxas models CC.CF as the hardware carry, so `subc` after `add.cc` sees the inverted flag.
// Every chain below stays inside one family (add.cc -> addc, sub.cc -> subc).

// x * 2^-64 mod p for x = hi:lo.  m = lo * (2^32 + 1) mod 2^64 = {a1, l0} with (e, a1) = l1 + l0;
// q = (m * p) >> 64 = m - a1 - e;  r = hi - q (+ p on borrow).  Any u64 in, u64 out; canonical out when hi < p.
GL_D u64 gl_mont_reduce(u64 hi, u64 lo) {
    const u32 l0 = (u32)lo, l1 = (u32)(lo >> 32), h0 = (u32)hi, h1 = (u32)(hi >> 32);
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 a1, ae, b0, b1, m, t0, t1;\n\t"
        "add.cc.u32  a1, %3, %2;\n\t"      // a1 = l1 + l0, CF = e
        "addc.u32    ae, a1, 0;\n\t"       // a1 + e (never wraps: l0 + l1 <= 2^33 - 2)
        "sub.cc.u32  b0, %2, ae;\n\t"      // q = {a1, l0} - (a1 + e)
        "subc.u32    b1, a1, 0;\n\t"
        "sub.cc.u32  t0, %4, b0;\n\t"      // r = hi - q, CF = borrow
        "subc.cc.u32 t1, %5, b1;\n\t"
        "subc.u32    m, 0, 0;\n\t"         // m = borrow ? 0xFFFFFFFF : 0
        "sub.cc.u32  %0, t0, m;\n\t"       // r -= borrow * EPS  (== r + p mod 2^64)
        "subc.u32    %1, t1, 0;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(l0), "r"(l1), "r"(h0), "r"(h1));
    return ((u64)r1 << 32) | r0;
}
GL_D u64 gl_mmul(u64 a, u64 b) { return gl_mont_reduce(__umul64hi(a, b), a * b); }   // a * b * 2^-64
GL_D u64 gl_msqr(u64 a) { return gl_mmul(a, a); }
// x * 2^64 mod p = x0 * EPS - x1 (x = x1 * 2^32 + x0): 1 IMAD.WIDE + 5 ALU.  Any u64 in, u64 out.
GL_D u64 gl_to_mont(u64 x) {
    const u64 t = (u64)(u32)x * 0xFFFFFFFFu;
    const u32 t0 = (u32)t, t1 = (u32)(t >> 32), x1 = (u32)(x >> 32);
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 s0, s1, m;\n\t"
        "sub.cc.u32  s0, %2, %4;\n\t"
        "subc.cc.u32 s1, %3, 0;\n\t"
        "subc.u32    m, 0, 0;\n\t"
        "sub.cc.u32  %0, s0, m;\n\t"
        "subc.u32    %1, s1, 0;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(t0), "r"(t1), "r"(x1));
    return ((u64)r1 << 32) | r0;
}
GL_D u64 gl_from_mont(u64 x) { return gl_mont_reduce(0, x); }   // canonical
#define GL_MONT_ONE 0xFFFFFFFFULL   // 1 * 2^64 mod p

// a + t with t <= p - 1 (a any u64): the wrapped sum is < t, so adding EPS once cannot wrap again.  3 ALU + 1 IMAD.WIDE.
GL_D u64 gl_addc(u64 a, u64 t) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), t0 = (u32)t, t1 = (u32)(t >> 32);
    u32 s0, s1, c;
    asm("add.cc.u32  %0, %3, %5;\n\t"
        "addc.cc.u32 %1, %4, %6;\n\t"
        "addc.u32    %2, 0, 0;"
        : "=r"(s0), "=r"(s1), "=r"(c)
        : "r"(a0), "r"(a1), "r"(t0), "r"(t1));
    return (u64)c * 0xFFFFFFFFu + (((u64)s1 << 32) | s0);
}
// a - t with t <= p - 1 (a any u64).  5 ALU.
GL_D u64 gl_subc(u64 a, u64 t) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), t0 = (u32)t, t1 = (u32)(t >> 32);
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 s0, s1, m;\n\t"
        "sub.cc.u32  s0, %2, %4;\n\t"
        "subc.cc.u32 s1, %3, %5;\n\t"
        "subc.u32    m, 0, 0;\n\t"
        "sub.cc.u32  %0, s0, m;\n\t"
        "subc.u32    %1, s1, 0;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(a0), "r"(a1), "r"(t0), "r"(t1));
    return ((u64)r1 << 32) | r0;
}

GL_D u64 gl_pow(u64 a, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = gl_mul(r, a);
        a = gl_mul(a, a);
        e >>= 1;
    }
    return r;
}
GL_D u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }

// ---- cubic extension (f3g.js:94-102): elements are 3 consecutive u64 ------------------------------------
struct gl3 {
    u64 c[3];
};

GL_D gl3 gl3_add(const gl3& a, const gl3& b) { return gl3{{gl_add(a.c[0], b.c[0]), gl_add(a.c[1], b.c[1]), gl_add(a.c[2], b.c[2])}}; }
GL_D gl3 gl3_sub(const gl3& a, const gl3& b) { return gl3{{gl_sub(a.c[0], b.c[0]), gl_sub(a.c[1], b.c[1]), gl_sub(a.c[2], b.c[2])}}; }
GL_D gl3 gl3_scale(const gl3& a, u64 s) { return gl3{{gl_mul(a.c[0], s), gl_mul(a.c[1], s), gl_mul(a.c[2], s)}}; }
// s_mont = s * 2^64 (canonical): plain product a * s
GL_D gl3 gl3_mscale(const gl3& a, u64 s_mont) { return gl3{{gl_mmul(a.c[0], s_mont), gl_mmul(a.c[1], s_mont), gl_mmul(a.c[2], s_mont)}}; }
GL_D gl3 gl3_canon(const gl3& a) { return gl3{{gl_canon(a.c[0]), gl_canon(a.c[1]), gl_canon(a.c[2])}}; }
// Lazy accumulator of 64 x 64 products: three 64-bit columns (weights 1, 2^32, 2^64) with their carry counts.  A product costs
// 4 IMAD.WIDE with carry-out + the carry adds (ptxas fuses mad.lo.cc / madc.hi.cc and merges the carries of neighbouring products).
struct GlAcc {
    u32 a0l, a0h, c0, a1l, a1h, c1, a2l, a2h, c2;
};
GL_D void gl_acc_madc(u32& lo, u32& hi, u32& c, u32 x, u32 y) {
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(lo), "+r"(hi), "+r"(c) : "r"(x), "r"(y));
}
GL_D void gl_acc_mac(GlAcc& A, u64 x, u64 y) {       // A += x * y, any u64 representatives
    const u32 x0 = (u32)x, x1 = (u32)(x >> 32), y0 = (u32)y, y1 = (u32)(y >> 32);
    gl_acc_madc(A.a0l, A.a0h, A.c0, x0, y0);
    gl_acc_madc(A.a1l, A.a1h, A.c1, x0, y1);
    gl_acc_madc(A.a1l, A.a1h, A.c1, x1, y0);
    gl_acc_madc(A.a2l, A.a2h, A.c2, x1, y1);
}
// a0 + c0 2^64 + (a1 + c1 2^64) 2^32 + (a2 + c2 2^64) 2^64 mod p for up to ~2^28 accumulated products: the low 128 bits go through
// gl_reduce128, what lies above 2^128 (top < 2^32) comes off as top * 2^32 (2^128 = -2^32 mod p) -- folded with 2^64 = 2^32 - 1 first, so
// that the subtrahend is canonical.
GL_D u64 gl_acc_reduce(const GlAcc& A) {
    u32 l1, h0, h1, top;
    asm("{\n\t"
        "add.cc.u32   %0, %4, %5;\n\t"        // a0h + a1l
        "addc.cc.u32  %1, %6, %7;\n\t"        // a2l + a1h + carry
        "addc.cc.u32  %2, %8, 0;\n\t"
        "addc.u32     %3, %9, 0;\n\t"         // c2 + carry
        "add.cc.u32   %1, %1, %10;\n\t"       // + c0
        "addc.cc.u32  %2, %2, %11;\n\t"       // + c1 (weight 2^96)
        "addc.u32     %3, %3, 0;\n\t"
        "}"
        : "=&r"(l1), "=&r"(h0), "=&r"(h1), "=&r"(top)
        : "r"(A.a0h), "r"(A.a1l), "r"(A.a2l), "r"(A.a1h), "r"(A.a2h), "r"(A.c2), "r"(A.c0), "r"(A.c1));
    const u64 r = gl_reduce128(((u64)h1 << 32) | h0, ((u64)l1 << 32) | A.a0l);
    return gl_subc(r, (u64)top << 32);             // top * 2^32 < p: canonical
}

// Product modulo x^3 - x - 1 (f3g.js:94-102): with c_k the coefficients of the plain polynomial product, x^3 = x + 1 and x^4 = x^2 + x give
//     r0 = c0 + c3,  r1 = c1 + c3 + c4,  r2 = c2 + c4.
// The 12 partial products are accumulated lazily (no reduction between them) and every coordinate is reduced once: ~170 instructions
// against ~300 for the Karatsuba form with a reduction per product (6 multiplications + 15 modular additions).  Any u64 in and out.
GL_D gl3 gl3_mul(const gl3& a, const gl3& b) {
    GlAcc r0 = {0, 0, 0, 0, 0, 0, 0, 0, 0}, r1 = r0, r2 = r0;
    gl_acc_mac(r0, a.c[0], b.c[0]);
    gl_acc_mac(r0, a.c[1], b.c[2]);
    gl_acc_mac(r0, a.c[2], b.c[1]);
    gl_acc_mac(r1, a.c[0], b.c[1]);
    gl_acc_mac(r1, a.c[1], b.c[0]);
    gl_acc_mac(r1, a.c[1], b.c[2]);
    gl_acc_mac(r1, a.c[2], b.c[1]);
    gl_acc_mac(r1, a.c[2], b.c[2]);
    gl_acc_mac(r2, a.c[0], b.c[2]);
    gl_acc_mac(r2, a.c[1], b.c[1]);
    gl_acc_mac(r2, a.c[2], b.c[0]);
    gl_acc_mac(r2, a.c[2], b.c[2]);
    gl3 r;
    r.c[0] = gl_acc_reduce(r0);
    r.c[1] = gl_acc_reduce(r1);
    r.c[2] = gl_acc_reduce(r2);
    return r;
}

// ---- host-side helpers (setup constants only: twiddle seeds, n^-1, shift powers) -------------------------
static inline u64 glh_mul(u64 a, u64 b) {
    unsigned __int128 x = (unsigned __int128)a * b;
    return (u64)(x % GL_P);
}
static inline u64 glh_pow(u64 a, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = glh_mul(r, a);
        a = glh_mul(a, a);
        e >>= 1;
    }
    return r;
}
static inline u64 glh_inv(u64 a) { return glh_pow(a, GL_P - 2); }
static inline u64 glh_to_mont(u64 a) { return glh_mul(a % GL_P, 0xFFFFFFFFULL); }   // a * 2^64 mod p
static inline u64 glh_root(unsigned s) {  // w[s], order 2^s (fft.js:45-50)
    u64 w = GL_W32;
    for (unsigned i = 32; i > s; i--) w = glh_mul(w, w);
    return w;
}
