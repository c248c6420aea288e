// Drop-in for src/helpers/fft/fft_p.js (exports at :299-302): same names, same arguments, same buffer contract
// (caller allocates buffDst, buffSrc untouched, buffDst fully overwritten), BigBuffers of any size (multi-page included).
// The work runs in libpil2gpu.so on a worker thread; the returned Promise resolves when buffDst is complete.
"use strict";
const { addon, context, inPages, outPages } = require("./pil2gpu.js");

// fft_p.js:178-180 / :182-184 -> pil2gpu_ntt_paged
async function fft(buffSrc, nPols, nBits, buffDst) {
    const dst = outPages(buffDst);
    await addon.nttPaged(context(), inPages(buffSrc), dst.pages, nPols, nBits, 0);
    dst.commit();
}
async function ifft(buffSrc, nPols, nBits, buffDst) {
    const dst = outPages(buffDst);
    await addon.nttPaged(context(), inPages(buffSrc), dst.pages, nPols, nBits, 1);
    dst.commit();
}
// fft_p.js:187-297 -> pil2gpu_lde_paged
async function interpolate(buffSrc, nPols, nBits, buffDst, nBitsExt) {
    const dst = outPages(buffDst);
    await addon.ldePaged(context(), inPages(buffSrc), dst.pages, nPols, nBits, nBitsExt);
    dst.commit();
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
