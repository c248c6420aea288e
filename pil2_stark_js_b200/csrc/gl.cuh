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
GL_D gl3 gl3_canon(const gl3& a) { return gl3{{gl_canon(a.c[0]), gl_canon(a.c[1]), gl_canon(a.c[2])}}; }
GL_D gl3 gl3_mul(const gl3& a, const gl3& b) {
    // Karatsuba-style product modulo x^3 - x - 1, same operation count as f3g.js:94-102
    u64 A = gl_mul(gl_add(a.c[0], a.c[1]), gl_add(b.c[0], b.c[1]));
    u64 B = gl_mul(gl_add(a.c[0], a.c[2]), gl_add(b.c[0], b.c[2]));
    u64 C = gl_mul(gl_add(a.c[1], a.c[2]), gl_add(b.c[1], b.c[2]));
    u64 D = gl_mul(a.c[0], b.c[0]);
    u64 E = gl_mul(a.c[1], b.c[1]);
    u64 F = gl_mul(a.c[2], b.c[2]);
    u64 G = gl_sub(D, E);
    gl3 r;
    r.c[0] = gl_sub(gl_add(C, G), F);
    r.c[1] = gl_sub(gl_sub(gl_sub(gl_add(A, C), E), E), D);
    r.c[2] = gl_sub(B, G);
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
static inline u64 glh_root(unsigned s) {  // w[s], order 2^s (fft.js:45-50)
    u64 w = GL_W32;
    for (unsigned i = 32; i > s; i--) w = glh_mul(w, w);
    return w;
}
