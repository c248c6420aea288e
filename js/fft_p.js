// Drop-in for src/helpers/fft/fft_p.js (exports at :299-302): same names, same arguments, same buffer contract
// (caller allocates buffDst, buffSrc untouched, buffDst fully overwritten).  The work runs in libpil2gpu.so.
"use strict";
const { addon, context, pagesOf } = require("./pil2gpu.js");

// fft_p.js:178-180 / :182-184
async function fft(buffSrc, nPols, nBits, buffDst) {
    addon.nttPaged(context(), pagesOf(buffSrc), pagesOf(buffDst), nPols, nBits, 0);
}
async function ifft(buffSrc, nPols, nBits, buffDst) {
    addon.nttPaged(context(), pagesOf(buffSrc), pagesOf(buffDst), nPols, nBits, 1);
}
// fft_p.js:187-297 -> pil2gpu_lde_paged
async function interpolate(buffSrc, nPols, nBits, buffDst, nBitsExt) {
    addon.ldePaged(context(), pagesOf(buffSrc), pagesOf(buffDst), nPols, nBits, nBitsExt);
}
// fft_p.js:20-32: a pure row permutation on caller buffers; kept in JS (it is not on the accelerated path).
function traspose(buffDst, buffSrc, nPols, nBits, trasposeBits) {
    const n = 1 << nBits, w = 1 << trasposeBits, h = n / w;
    for (let i = 0; i < w; i++) {
        for (let j = 0; j < h; j++) {
            const fi = j * w + i, di = i * h + j;
            buffDst.set(buffSrc.slice(fi * nPols, fi * nPols + nPols), di * nPols);
        }
    }
}
module.exports.fft = fft;
module.exports.ifft = ifft;
module.exports.interpolate = interpolate;
module.exports.traspose = traspose;
