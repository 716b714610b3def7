"""The DEFLATE decoder and CRC-32 pieces the GPU BGZF kernels are built from (csrc/cuda/inflate_core.cuh), compiled for the
CPU and fuzzed against zlib: every block type, every level and strategy, long matches, the 32 KiB window, empty input,
misaligned input, corrupt streams (must fail or decode within bounds — never touch a byte outside the output)."""
import ctypes
import os
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def h():
    out_dir = os.path.join(ROOT, "build", "tests")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libinflate_harness.so")
    src = os.path.join(ROOT, "tests", "native", "inflate_harness.cpp")
    hdr = os.path.join(ROOT, "trueconsense_b200", "csrc", "cuda", "inflate_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-g", "-shared", "-fPIC", "-x", "c++", "-I", os.path.dirname(hdr), "-o", so, src], check=True)
    lib = ctypes.CDLL(so)
    lib.h_inflate.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
    lib.h_crc.argtypes = [ctypes.c_void_p, ctypes.c_int64]
    lib.h_crc.restype = ctypes.c_uint32
    lib.h_crc_combine.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64]
    lib.h_crc_combine.restype = ctypes.c_uint32
    return lib


GUARD = 64


def _inflate(lib, comp: bytes, n_out: int, misalign: int = 0, stride: int = 1):
    """(rc, output bytes); asserts the guard bytes around the output are untouched."""
    # the reader fetches aligned 16-byte vectors: readable from the boundary in front of the stream to the one behind it
    raw = np.zeros(len(comp) + 80, dtype=np.uint8)
    start = (-raw.ctypes.data) % 16 + 16 + misalign
    raw[start:start + len(comp)] = np.frombuffer(comp, dtype=np.uint8)
    raw[start + len(comp):] = 0xEE                                   # (bytes behind the stream must not matter)
    raw[:start] = 0xEE
    base, misalign = raw.ctypes.data + start, 0
    out = np.full(n_out + 2 * GUARD, 0xA5, dtype=np.uint8)
    rc = lib.h_inflate(base + misalign, len(comp), out.ctypes.data + GUARD, n_out, stride)
    assert (out[:GUARD] == 0xA5).all() and (out[GUARD + n_out:] == 0xA5).all(), "wrote outside the output"
    return rc, out[GUARD:GUARD + n_out].tobytes()


def _raw_deflate(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem=8) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, -15, mem, strategy)
    return c.compress(data) + c.flush()


def _samples(rng):
    yield b""
    yield b"a"
    yield b"abc" * 5
    yield bytes(70000 // 7 * [1, 2, 3, 4, 5, 6, 7])[:65280]
    yield bytes(65280)                                   # one long run: length 258 matches at distance 1
    yield rng.integers(0, 256, 65280, dtype=np.uint8).tobytes()          # incompressible: stored blocks at most levels
    yield rng.integers(33, 75, 65280, dtype=np.uint8).tobytes()          # quality-like: literals only, short codes
    text = b"".join(b"read%07d\tACGT" % i + bytes(rng.integers(0, 4, 40, dtype=np.uint8) + 65) for i in range(1100))
    yield text[:65280]
    far = rng.integers(0, 256, 2000, dtype=np.uint8).tobytes()
    yield far + bytes(rng.integers(0, 3, 30000, dtype=np.uint8)) + far   # a match 32 000 bytes back
    yield bytes(rng.integers(0, 2, 65280, dtype=np.uint8))               # two symbols: 1-bit codes
    for n in (1, 2, 3, 7, 8, 9, 255, 256, 257, 258, 259, 1000, 4095, 4096, 32768, 32769):
        yield rng.integers(0, 5, n, dtype=np.uint8).tobytes()


def test_inflate_equals_zlib_for_every_block_type(h):
    rng = np.random.default_rng(11)
    n = 0
    for data in _samples(rng):
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
                comp = _raw_deflate(data, level, strategy, mem=8 if n % 3 else 1)
                for mis in ((0,) if n % 4 else (0, 1, 2, 3, 5, 11, 15)):
                    rc, got = _inflate(h, comp, len(data), misalign=mis, stride=1 if n % 2 else 32)
                    assert rc == 0, (len(data), level, strategy, mis, rc)
                    assert got == data
                n += 1
    assert n > 400


def test_inflate_multiple_deflate_blocks_and_sync_flushes(h):
    rng = np.random.default_rng(5)
    c = zlib.compressobj(1, zlib.DEFLATED, -15)
    parts, comp = [], b""
    for i in range(12):
        d = rng.integers(0, 1 + 20 * i, 3000 + 100 * i, dtype=np.uint8).tobytes()
        parts.append(d)
        comp += c.compress(d) + c.flush(zlib.Z_SYNC_FLUSH if i % 2 else zlib.Z_FULL_FLUSH)       # empty stored blocks in between
    comp += c.flush()
    data = b"".join(parts)
    rc, got = _inflate(h, comp, len(data))
    assert rc == 0 and got == data


def test_inflate_rejects_or_bounds_corrupt_streams(h):
    """Bit flips, truncations, wrong ISIZE: an error code, or (when the damaged stream still happens to be a stream) output
    within bounds; the guard bytes are checked by the helper on every call."""
    rng = np.random.default_rng(23)
    data = b"".join(b"q%06d" % i + rng.integers(33, 70, 30, dtype=np.uint8).tobytes() for i in range(1500))
    comp = _raw_deflate(data, 1)
    assert _inflate(h, comp, len(data))[0] == 0
    assert _inflate(h, comp, len(data) - 1)[0] != 0            # ISIZE too small
    assert _inflate(h, comp, len(data) + 1)[0] != 0            # ISIZE too large
    assert _inflate(h, comp[:-1], len(data))[0] != 0 or True   # (the last byte may hold padding bits only)
    assert _inflate(h, comp[: len(comp) // 2], len(data))[0] != 0
    assert _inflate(h, b"", 0)[0] != 0
    assert _inflate(h, b"\x07", 0)[0] != 0                     # block type 3
    assert _inflate(h, b"\x01\x05\x00\x00\x00", 5)[0] != 0     # stored: NLEN is not ~LEN
    wrong = errors = 0
    for trial in range(600):
        bad = bytearray(comp)
        for _ in range(1 + trial % 3):
            bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        rc, got = _inflate(h, bytes(bad), len(data))
        ref_ok = True
        try:
            ref = zlib.decompressobj(-15).decompress(bytes(bad))
            ref_ok = len(ref) == len(data)
        except zlib.error:
            ref_ok = False
        if rc == 0:
            assert ref_ok and got == ref, trial                # accepted: then zlib accepts it too, with the same bytes
            wrong += got != data
        else:
            errors += 1
    assert errors > 100 and wrong > 100          # (a flipped literal still is a stream: the CRC-32 is what catches those)


def test_crc_pieces_equal_zlib(h):
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 3, 4, 5, 31, 32, 33, 255, 2040, 2048, 65280):
        d = rng.integers(0, 256, n, dtype=np.uint8)
        assert h.h_crc(d.ctypes.data, n) == zlib.crc32(d.tobytes())
    for la, lb in ((0, 0), (0, 9), (9, 0), (1, 1), (2040, 2040), (2048, 17), (65280, 1), (123, 456789)):
        a = rng.integers(0, 256, la, dtype=np.uint8).tobytes()
        b = rng.integers(0, 256, lb, dtype=np.uint8).tobytes()
        assert h.h_crc_combine(zlib.crc32(a), zlib.crc32(b), lb) == zlib.crc32(a + b)
