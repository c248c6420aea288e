"""Mirror of src/stark/stark_constsPolsFile.js: the `.cnts` setup artefact (fixed-polynomial evaluations, the const tree,
x_n and x_ext), plus a loader that sends the const tree straight into a device-resident tree handle.

Container = the iden3 "binfile" framing the reference gets from @iden3/binfileutils 0.0.11 (package.json; the dependency is
not vendored in the reference tree, its published layout is restated here):
    bytes 0..3   type, ASCII "cnts"                                   (createBinFile, stark_constsPolsFile.js:22)
    u32 LE       version (1)
    u32 LE       number of sections (CONSTS_PS_NSECTIONS = 5, stark_constsPols_constants.js:2)
    per section: u32 LE section id, u64 LE payload size, payload       (startWriteSection / endWriteSection)
Payloads (all u64 words little-endian, lengths as u32 LE like the reference's writeULE32):
    2  fixed pols evals   u32 n, n words                               (:44-50)
    3  const tree         u32 width, u32 height, u32 nElements, elements, u32 nNodes, nodes   (:52-65)
    4  x_n                u32 n, n words                               (:68-75)
    5  x_ext              u32 n, n words                               (:78-83)
(The reference's u32 length fields cap a section at 2^32 - 1 words; larger artefacts do not fit this format there either.)
"""
import struct

import numpy as np

CONSTS_PS_NSECTIONS = 5                      # stark_constsPols_constants.js:2-8
CONSTS_PS_CONST_POLS_EVALS_SECTION = 2
CONSTS_PS_CONST_TREE_SECTION = 3
CONSTS_PS_X_N_SECTION = 4
CONSTS_PS_X_EXT_SECTION = 5


def _words(a):
    return np.ascontiguousarray(a, dtype="<u8").reshape(-1)


def _section(f, sid, parts):
    f.write(struct.pack("<I", sid))
    pos = f.tell()
    f.write(struct.pack("<Q", 0))
    start = f.tell()
    for p in parts:
        if isinstance(p, (bytes, bytearray)):
            f.write(p)
        else:
            p.tofile(f)
    end = f.tell()
    f.seek(pos)
    f.write(struct.pack("<Q", end - start))
    f.seek(end)


def writePilStarkConstsFile(consts, constsFilename):
    """stark_constsPolsFile.js:18-42.  consts = {"fixedPolsEvals", "constTree": {"width","height","elements","nodes"}, "x_n", "x_ext"}."""
    t = consts["constTree"]
    for name, a in (("fixedPolsEvals", consts["fixedPolsEvals"]), ("elements", t["elements"]), ("nodes", t["nodes"]), ("x_n", consts["x_n"]),
                    ("x_ext", consts["x_ext"])):
        if _words(a).size >= 1 << 32:
            raise ValueError(f"{name}: {_words(a).size} words do not fit the format's u32 length field")
    with open(constsFilename, "wb") as f:
        f.write(b"cnts")
        f.write(struct.pack("<II", 1, CONSTS_PS_NSECTIONS))
        ev = _words(consts["fixedPolsEvals"])
        _section(f, CONSTS_PS_CONST_POLS_EVALS_SECTION, [struct.pack("<I", ev.size), ev])
        el, nd = _words(t["elements"]), _words(t["nodes"])
        _section(f, CONSTS_PS_CONST_TREE_SECTION, [struct.pack("<II", t["width"], t["height"]), struct.pack("<I", el.size), el,
                                                   struct.pack("<I", nd.size), nd])
        xn, xe = _words(consts["x_n"]), _words(consts["x_ext"])
        _section(f, CONSTS_PS_X_N_SECTION, [struct.pack("<I", xn.size), xn])
        _section(f, CONSTS_PS_X_EXT_SECTION, [struct.pack("<I", xe.size), xe])


def _read_sections(f):
    if f.read(4) != b"cnts":
        raise ValueError("Invalid File format")                       # readBinFile's check
    version, n_sections = struct.unpack("<II", f.read(8))
    if version > 1:
        raise ValueError("Version not supported")
    sections = {}
    for _ in range(n_sections):
        hdr = f.read(12)
        if len(hdr) < 12:
            break
        sid, size = struct.unpack("<IQ", hdr)
        sections.setdefault(sid, []).append((f.tell(), size))
        f.seek(size, 1)
    return sections


def _unique(sections, sid):
    if sid not in sections:
        raise ValueError(f"Missing section {sid}")
    if len(sections[sid]) > 1:
        raise ValueError(f"Section Duplicated {sid}")                  # startReadUniqueSection
    return sections[sid][0]


def _read_words(f, n):
    a = np.fromfile(f, dtype="<u8", count=n)
    if a.size != n:
        raise ValueError("consts file is truncated")
    return a.astype(np.uint64, copy=False)


def readPilStarkConstsFile(constsFilename, ctx=None, tree_to_device=False):
    """stark_constsPolsFile.js:96-131.  With tree_to_device the const tree section goes straight into a DeviceTree (streamed in
    chunks, no re-hashing) and consts["constTree"] is that handle; MerkleHash.getGroupProof / root accept it."""
    consts = {}
    with open(constsFilename, "rb") as f:
        sections = _read_sections(f)
        pos, _ = _unique(sections, CONSTS_PS_CONST_POLS_EVALS_SECTION)
        f.seek(pos)
        (n,) = struct.unpack("<I", f.read(4))
        consts["fixedPolsEvals"] = _read_words(f, n)
        pos, _ = _unique(sections, CONSTS_PS_CONST_TREE_SECTION)
        f.seek(pos)
        width, height, n_el = struct.unpack("<III", f.read(12))
        if tree_to_device:
            from .context import default_context
            g = ctx or default_context()
            if n_el != width * height:
                raise ValueError(f"const tree section: {n_el} elements, expected width*height = {width * height}")
            tree = g.tree_alloc(width, height)
            chunk = 1 << 25
            for which, total in ((0, n_el), (1, None)):
                if which == 1:
                    (total,) = struct.unpack("<I", f.read(4))
                    if total != g.merkle_nnodes(height):       # a short section would leave uninitialised device words in a tree
                        tree.free()                            # that later serves proofs
                        raise ValueError(f"const tree section: {total} node words, expected {g.merkle_nnodes(height)}")
                off = 0
                while off < total:
                    m = min(chunk, total - off)
                    tree.fill(which, off, _read_words(f, m))
                    off += m
            consts["constTree"] = tree
        else:
            el = _read_words(f, n_el)
            (n_nd,) = struct.unpack("<I", f.read(4))
            consts["constTree"] = {"width": width, "height": height, "elements": el, "nodes": _read_words(f, n_nd)}
        for key, sid in (("x_n", CONSTS_PS_X_N_SECTION), ("x_ext", CONSTS_PS_X_EXT_SECTION)):
            pos, _ = _unique(sections, sid)
            f.seek(pos)
            (n,) = struct.unpack("<I", f.read(4))
            consts[key] = _read_words(f, n)
    return consts
