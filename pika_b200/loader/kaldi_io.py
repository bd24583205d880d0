"""Native readers for the on-disk formats of the reference recipes (no PyKaldi):
  * Kaldi int-vector archives ``ark:path`` / ``ark,t:path`` (labels; loader/otf_utt_loader.py:209 SequentialIntVectorReader)
  * Kaldi text double matrix (global CMVN stats; trainer/train_transducer_bmuf_otfaug.py:341-346)
  * ``.mrk`` / ``.seq`` raw-PCM shards (utils/wav_to_seq.py:36-38) and ``.lst`` triplets (loader/otf_utt_loader.py:124-129)
"""
import struct

import numpy as np


def parse_rspecifier(rspec):
    if ":" in rspec and rspec.split(":", 1)[0].replace(",", "").replace("ark", "").replace("t", "").replace("scp", "") == "":
        kind, path = rspec.split(":", 1)
        if "scp" in kind:
            raise NotImplementedError("scp rspecifiers are not used by the recipes")
        return path
    return rspec


def read_int_vector_ark(rspec):
    """Yields (uttid, list[int]) in file order; text or binary Kaldi archives."""
    path = parse_rspecifier(rspec)
    with open(path, "rb") as f:
        data = f.read()
    pos, n = 0, len(data)
    while pos < n:
        while pos < n and data[pos:pos + 1] in b" \n\t\r":
            pos += 1
        if pos >= n:
            break
        sp = pos
        while data[sp:sp + 1] not in b" \n\t":
            sp += 1
        key = data[pos:sp].decode()
        pos = sp + 1
        if data[pos:pos + 2] == b"\0B":                       # binary: \0B \4 <int32 n> (\4 <int32>)*
            pos += 2
            assert data[pos] == 4
            cnt = struct.unpack_from("<i", data, pos + 1)[0]
            pos += 5
            vals = []
            for _ in range(cnt):
                assert data[pos] == 4
                vals.append(struct.unpack_from("<i", data, pos + 1)[0])
                pos += 5
            yield key, vals
        else:                                                 # text: rest of the line
            nl = data.find(b"\n", pos)
            nl = n if nl < 0 else nl
            yield key, [int(t) for t in data[pos:nl].split()]
            pos = nl + 1


def read_kaldi_text_matrix(path):
    """``[ r0c0 r0c1 ...\\n r1c0 ... ]`` -> float64 ndarray (CMVN stats are 2 x (D+1))."""
    txt = open(path).read()
    lb, rb = txt.index("["), txt.rindex("]")
    rows = [r.split() for r in txt[lb + 1:rb].strip().split("\n") if r.strip()]
    return np.array([[float(v) for v in r] for r in rows], dtype=np.float64)


def cmvn_offset_scale(path, splice_width):
    """(offset, scale) float64 tiled lctx+1+rctx times (trainer/train_transducer_bmuf_otfaug.py:341-355)."""
    cmvn = read_kaldi_text_matrix(path)
    mean = cmvn[0][:-1] / cmvn[0][-1]
    var = cmvn[1][:-1] / cmvn[0][-1] - mean * mean
    if min(abs(var)) < 1.0e-20:
        raise ValueError("problematic cmvn_stats, variance too small")
    return np.tile(-mean, splice_width), np.tile(1.0 / np.sqrt(var), splice_width)


def read_lst(data_lst):
    out = []
    with open(data_lst, "r", encoding="utf-8") as f:
        for line in f:
            p = line.split()
            if p:
                out.append((p[0], p[1], p[2]))
    return out


def iter_mrk_seq(mrk_fn, seq_fn):
    """Yields (uttid, int16 ndarray) -- utils/wav_to_seq.py layout, odd byte counts truncated
    (loader/otf_utt_loader.py:213-217)."""
    with open(mrk_fn, "r", encoding="utf-8") as mrk, open(seq_fn, "rb") as seq:
        for line in mrk:
            p = line.split()
            seq.seek(int(p[1]))
            nb = int(p[2])
            nb -= nb % 2
            yield p[0], np.frombuffer(seq.read(nb), dtype="int16")


# ------------------------------------------------------------------------------------------------ float matrix tables
def _read_token(f):
    tok = b""
    while True:
        c = f.read(1)
        if not c:
            return None if not tok else tok.decode()
        if c in b" \n\t\r":
            if tok:
                return tok.decode()
            continue
        tok += c


def _read_compressed_matrix(f, fmt):
    """Kaldi CompressedMatrix payload after its token (matrix/compressed-matrix.h): float min_value, float range, int32 rows, int32 cols
    (the 'format' field of the global header is not written, the token carries it), then
      CM  (format 1, what make_fbank.sh / copy-feats --compress=true write): per column four uint16 percentiles (0, 25, 75, 100 % as
          min + range * v / 65535), then one byte per element, column-major; a byte decodes piecewise linearly between the percentiles
          (v <= 64 | 64 < v <= 192 | v > 192);
      CM2 (format 2): uint16 per element, row-major, min + range * v / 65535;   CM3 (format 3): uint8, min + range * v / 255."""
    vmin, vrange = struct.unpack("<ff", f.read(8))
    rows, cols = struct.unpack("<ii", f.read(8))
    if fmt == 1:
        pc = np.frombuffer(f.read(cols * 8), dtype=np.uint16).reshape(cols, 4).astype(np.float32)
        pc = np.float32(vmin) + np.float32(vrange) * np.float32(1.52590218966964e-05) * pc          # [cols, 4]: p0, p25, p75, p100
        d = np.frombuffer(f.read(cols * rows), dtype=np.uint8).reshape(cols, rows).astype(np.float32)
        p0, p25, p75, p100 = (pc[:, i:i + 1] for i in range(4))
        out = np.where(d <= 64, p0 + (p25 - p0) * d * np.float32(1 / 64.0),
                       np.where(d <= 192, p25 + (p75 - p25) * (d - 64) * np.float32(1 / 128.0),
                                p75 + (p100 - p75) * (d - 192) * np.float32(1 / 63.0)))
        return np.ascontiguousarray(out.T, dtype=np.float32)
    if fmt == 2:
        d = np.frombuffer(f.read(rows * cols * 2), dtype=np.uint16).reshape(rows, cols).astype(np.float32)
        return (np.float32(vmin) + np.float32(vrange) * np.float32(1.52590218966964e-05) * d).astype(np.float32)
    d = np.frombuffer(f.read(rows * cols), dtype=np.uint8).reshape(rows, cols).astype(np.float32)
    return (np.float32(vmin) + np.float32(vrange) * np.float32(1.0 / 255.0) * d).astype(np.float32)


def _read_binary_matrix(f):
    """after ``\0B``: 'FM ' | 'DM ' , \4 <int32 rows> \4 <int32 cols>, row-major data (Kaldi's uncompressed matrix layout), or a
    compressed matrix 'CM ' | 'CM2 ' | 'CM3 ' (the default of Kaldi's feature dumps)"""
    kind = f.read(3)
    if kind == b"CM ":
        return _read_compressed_matrix(f, 1)
    if kind in (b"CM2", b"CM3"):
        assert f.read(1) == b" "
        return _read_compressed_matrix(f, 2 if kind == b"CM2" else 3)
    if kind not in (b"FM ", b"DM "):
        raise NotImplementedError("Kaldi matrix type %r" % kind)
    assert f.read(1) == b"\4"
    rows = struct.unpack("<i", f.read(4))[0]
    assert f.read(1) == b"\4"
    cols = struct.unpack("<i", f.read(4))[0]
    dt = np.float32 if kind == b"FM " else np.float64
    return np.frombuffer(f.read(rows * cols * np.dtype(dt).itemsize), dtype=dt).reshape(rows, cols).astype(np.float32)


def _read_text_matrix(f):
    """`` [ r0c0 r0c1 ...\\n  r1c0 ... ]`` following a key"""
    rows, cur = [], []
    tok = _read_token(f)
    assert tok == "[", "expected '[' at the start of a text matrix"
    buf = b""
    while True:
        c = f.read(1)
        if not c or c == b"]":
            break
        buf += c
    for line in buf.decode().split("\n"):
        vals = line.split()
        if vals:
            rows.append([float(v) for v in vals])
    return np.asarray(rows, dtype=np.float32).reshape(len(rows), -1)


def read_float_matrix_table(rspec):
    """Yields (uttid, float32 [frames, dim]) in file order from ``ark:path``, ``ark,t:path`` or ``scp:path`` (lines ``key path[:offset]``)
    -- the native stand-in for kaldi.util.table.SequentialMatrixReader (loader/utt_loader.py:80,158)."""
    kind = rspec.split(":", 1)[0] if ":" in rspec else "ark"
    path = rspec.split(":", 1)[1] if ":" in rspec else rspec
    if "scp" in kind:
        for line in open(path):
            line = line.strip()
            if not line:
                continue
            key, loc = line.split(None, 1)
            fn, off = loc, None
            if ":" in loc and loc.rsplit(":", 1)[1].isdigit():
                fn, off = loc.rsplit(":", 1)
            with open(fn, "rb") as f:
                if off is not None:
                    f.seek(int(off))                           # offset of the matrix itself (what copy-feats writes into an scp)
                else:
                    _read_token(f)                             # a file holding "key <matrix>"
                pos = f.tell()
                if f.read(2) == b"\0B":
                    yield key, _read_binary_matrix(f)
                else:
                    f.seek(pos)
                    yield key, _read_text_matrix(f)
        return
    with open(path, "rb") as f:
        while True:
            key = _read_token(f)
            if key is None:
                break
            pos = f.tell()
            if f.read(2) == b"\0B":
                yield key, _read_binary_matrix(f)
            else:
                f.seek(pos)
                yield key, _read_text_matrix(f)


def write_float_matrix_ark(path, items, text=False):
    """(uttid, [frames, dim] array) pairs -> Kaldi archive (binary 'FM ' or text); returns {uttid: byte offset of the matrix} for an scp"""
    offs = {}
    with open(path, "wb") as f:
        for key, mat in items:
            mat = np.ascontiguousarray(mat, dtype=np.float32)
            f.write(key.encode() + b" ")
            offs[key] = f.tell()
            if text:
                f.write(b" [\n")
                for r in mat:
                    f.write(("  " + " ".join("%.7g" % v for v in r) + "\n").encode())
                f.seek(f.tell() - 1)
                f.write(b" ]\n")
            else:
                f.write(b"\0BFM \4" + struct.pack("<i", mat.shape[0]) + b"\4" + struct.pack("<i", mat.shape[1]) + mat.tobytes())
    return offs
