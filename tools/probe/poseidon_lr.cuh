// Probe-only reference copy of the permutation shipped BEFORE the limb-resident partial rounds (u64 state at every round
// boundary: S-box -> split -> MDS -> join for all 12 lanes in all 30 rounds), kept so that tools/probe/gl_probe.cu can
// still measure the old against the new.
#pragma once
#include "poseidon.cuh"

GL_D void poseidon_mds(u64 x[12], const u32* __restrict__ rc_limbs) {
    u32 a[12], b[12], c[12];
#pragma unroll
    for (int j = 0; j < 12; j++) poseidon_split(x[j], a[j], b[j], c[j]);
    u32 ya[12], yb[12], yc[12];
    poseidon_mds_limb(ya, a, rc_limbs);
    poseidon_mds_limb(yb, b, rc_limbs + 1);
    poseidon_mds_limb(yc, c, rc_limbs + 2);
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = poseidon_join(ya[i], yb[i], yc[i]);
}
GL_D void poseidon_permute_mont_u64state(u64 x[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = gl_addc(x[i], POSEIDON_RC0[i]);
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        x[0] = poseidon_sbox(x[0]);
        if (r < 4 || r >= 26) {
#pragma unroll
            for (int i = 1; i < 12; i++) x[i] = poseidon_sbox(x[i]);
        }
        poseidon_mds(x, POSEIDON_RC_LIMBS + r * 36);
    }
}
