// Candidate P4: Montgomery S-box on the integer pipes + MDS on the FP64 pipe (exact: every intermediate is an integer of
// magnitude < 2^46).  B200 (sm_100a) issues DFMA/DADD at 16 lanes/clk/SMSP on a pipe of its own, idle otherwise.
#pragma once
#include "poseidon_mont.cuh"

__constant__ double POSEIDON_RC_F64[31 * 24] = {   // [round-1][lane][lo,hi]: 2^52 + half-word of the Montgomery-form constant
#include "poseidon_rc_f64.inc"
};

GL_D void poseidon_mds_fp64(u64 x[12], const double* __restrict__ rc) {
    const double MAGIC = 4503599627370496.0;   // 2^52
    double lo[12], hi[12], yl[12], yh[12];
#pragma unroll
    for (int j = 0; j < 12; j++) {
        lo[j] = __hiloint2double(0x43300000, (int)(u32)x[j]) - MAGIC;
        hi[j] = __hiloint2double(0x43300000, (int)(u32)(x[j] >> 32)) - MAGIC;
    }
    poseidon_mds_lanes<double>(yl, lo);
    poseidon_mds_lanes<double>(yh, hi);
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const double tl = yl[i] + rc[2 * i], th = yh[i] + rc[2 * i + 1];      // + 2^52 + rc half: integer now in the mantissa
        const u64 L = ((u64)((u32)__double2hiint(tl) & 0xFFFFFu) << 32) | (u32)__double2loint(tl);
        const u64 H = ((u64)((u32)__double2hiint(th) & 0xFFFFFu) << 32) | (u32)__double2loint(th);
        x[i] = poseidon_join64(L, H);
    }
}

GL_D void poseidon_permute_fp64(u64 x[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = gl_addc(x[i], POSEIDON_RC_MONT[i]);
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        x[0] = poseidon_sbox_mont(x[0]);
        if (r < 4 || r >= 26) {
#pragma unroll
            for (int i = 1; i < 12; i++) x[i] = poseidon_sbox_mont(x[i]);
        }
        poseidon_mds_fp64(x, POSEIDON_RC_F64 + 24 * r);
    }
}
