"""B200-native commit-phase hot path of pil2-stark-js: Goldilocks NTT/LDE, Poseidon-GL Merkle trees, FRI folding.

Python mirror of the three reference modules (same names, arguments and error behaviour):
    fft_p         fft / ifft / interpolate                 src/helpers/fft/fft_p.js
    merklehash_p  buildMerkleHash -> MerkleHash            src/helpers/hash/merklehash/merklehash_p.js
    fri           class FRI (fold / proofQueries / verify) src/stark/fri.js
    transcript    class Transcript                         src/helpers/transcript/transcript.js
    stark_gen_helpers  extendAndMerkelize, computeQStark, computeEvalsStark, ...   src/stark/stark_gen_helpers.js
    stark_consts_file  read/writePilStarkConstsFile (`.cnts`)                      src/stark/stark_constsPolsFile.js
    prover_helpers     callCalculateExps / calculateExps (constraint expressions)   src/prover/prover_helpers.js
All arithmetic runs in libpil2gpu.so (hand-written sm_100a CUDA behind the C ABI of include/pil2gpu.h); there is no
CPU fallback, so importing works anywhere but every operation needs a CUDA device.
"""
from ._lib import Pil2GpuError, OutOfRange, load as load_library   # noqa: F401
from .context import Context, DeviceTree, default_context          # noqa: F401
from .bigbuffer import BigBuffer                                    # noqa: F401
from . import fft_p, merklehash_p, fri, transcript, stark_gen_helpers, stark_consts_file, prover_helpers   # noqa: F401
from .fft_p import fft, ifft, interpolate                           # noqa: F401
from .merklehash_p import buildMerkleHash, MerkleHash               # noqa: F401
from .fri import FRI                                                # noqa: F401
from .transcript import Transcript                                  # noqa: F401
