// Candidate arithmetic for the probe: Montgomery-form Goldilocks multiply (R = 2^64) + cheap add/sub with one
// canonical operand.  See tools/probe/gl_probe.cu.
#pragma once
#include "gl.cuh"

// x * 2^-64 mod p for x = hi:lo.  Result is a 64-bit representative; it is canonical (< p) whenever hi < p, which holds
// for a product of any u64 with a canonical value.  8 ALU-pipe instructions, no multiply.
GL_D u64 gl_mont_reduce(u64 hi, u64 lo) {
    // m = lo * (2^32 + 1) mod 2^64 = {a1, l0} with (e, a1) = l1 + l0;  q = m - a1 - e = (m * p) >> 64;  r = hi - q (+ p on borrow)
    // Only same-family carry chains are used (add.cc -> addc, sub.cc -> subc): ptxas models CC.CF as the hardware carry,
    // so a subc after an add.cc would see the inverted flag.
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
GL_D u64 gl_mmul(u64 a, u64 b) { return gl_mont_reduce(__umul64hi(a, b), a * b); }
GL_D u64 gl_msqr(u64 a) { return gl_mmul(a, a); }

// a + t with t <= p - 1 (a any u64): one conditional correction, never a second wrap.
GL_D u64 gl_addc(u64 a, u64 t) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), t0 = (u32)t, t1 = (u32)(t >> 32);
    u32 s0, s1, c;
    asm("add.cc.u32  %0, %3, %5;\n\t"
        "addc.cc.u32 %1, %4, %6;\n\t"
        "addc.u32    %2, 0, 0;"            // c = carry
        : "=r"(s0), "=r"(s1), "=r"(c)
        : "r"(a0), "r"(a1), "r"(t0), "r"(t1));
    // += carry * EPS on the FMA pipe (one IMAD.WIDE); cannot wrap again because the wrapped sum is < t
    return (u64)c * 0xFFFFFFFFu + (((u64)s1 << 32) | s0);
}
// a - t with t <= p - 1 (a any u64).
GL_D u64 gl_subc(u64 a, u64 t) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), t0 = (u32)t, t1 = (u32)(t >> 32);
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 s0, s1, m;\n\t"
        "sub.cc.u32  s0, %2, %4;\n\t"
        "subc.cc.u32 s1, %3, %5;\n\t"
        "subc.u32    m, 0, 0;\n\t"         // m = borrow ? 0xFFFFFFFF : 0
        "sub.cc.u32  %0, s0, m;\n\t"       // -= borrow * EPS
        "subc.u32    %1, s1, 0;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(a0), "r"(a1), "r"(t0), "r"(t1));
    return ((u64)r1 << 32) | r0;
}
