"""K0: StringLookup on the device (csrc/string_lookup.cu) against the reference semantics
keras.layers.StringLookup(vocabulary=v, mask_token=None) -- v[i] -> i + 1, anything else -> 0 (src/models.py:70,73) --
as oracle.string_lookup states them (and as the reference run in tests/golden/reference_run.* exercises them).

Integer / byte work: every comparison is exact.  Without a GPU the product's per-id routine is compiled for the host by
tests/native/string_lookup_check.cu (the same lookup_one / build_table code, outside the product library) and run on the
same cases; with a GPU (-m gpu) the kernel itself runs through the C ABI and through the model call.
"""

import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle.models import make_vocab_index, string_lookup

ROOT = Path(__file__).resolve().parent.parent
NATIVE = ROOT / "tests" / "native"


def _cases():
    """(name, vocabulary, ids) with every edge the byte layout has."""
    rng = np.random.default_rng(7)
    ml = sorted(str(i) for i in range(1, 6041))                       # the reference's vocabularies: sorted numeric strings
    ids_ml = [ml[i] for i in rng.integers(0, len(ml), 3000)] + ["0", "6041", "", "nobody", "12345678901234567", "1 "]
    words = ["", "a", "ab", "abc", "b" * 15, "c" * 16, "d" * 17, "e" * 31, "f" * 32, "g" * 33, "héllo", "日本語", "x y", "A", "aB"]
    ids_w = words + ["b" * 14, "b" * 16, "c" * 15, "c" * 17, "d" * 16, "d" * 18, "e" * 32, "f" * 31, "g" * 34, "g" * 48,
                     "hello", "hé", "日本", "x", "x y ", "aa", "ABC", "é"]
    big = [f"user_{i:07d}_{(i * 2654435761) % 97}" for i in range(50000)]     # 50k words: long probe chains get exercised
    ids_big = [big[i] for i in rng.integers(0, len(big), 5000)] + [f"user_{i:07d}_x" for i in range(200)]
    dup = ["p", "q", "p", "r", "q", "p"]                              # a repeated word keeps its LAST position (dict order)
    return [("movielens", ml, ids_ml), ("edges", words, ids_w), ("50k", big, ids_big), ("dup", dup, ["p", "q", "r", "s"]),
            ("empty-vocab", [], ["a", ""]), ("empty-ids", ["a"], [])]


def _want(vocab, ids):
    return string_lookup(make_vocab_index(list(vocab)), ids)


@pytest.fixture(scope="module", params=["tags32", "tags2"])
def host_check(request):
    """The product's lookup routine compiled for the host: once as shipped, once with 2-bit tags (different words then share
    tags all the time, so the one-shot candidate is often the wrong word and the sequential byte-checked walk takes over)."""
    out = NATIVE / "_build" / f"libstring_lookup_check_{request.param}.so"
    src = NATIVE / "string_lookup_check.cu"
    prod = ROOT / "recommendation-system-maang-nvidia-_b200" / "csrc" / "string_lookup.cu"
    if not out.exists() or out.stat().st_mtime < max(src.stat().st_mtime, prod.stat().st_mtime):
        nvcc = "/usr/local/cuda/bin/nvcc"
        if not Path(nvcc).exists():
            pytest.skip("nvcc is not available to build the host check")
        out.parent.mkdir(exist_ok=True)
        extra = ["-DRS_STRING_LOOKUP_TAG_MASK=0x3u"] if request.param == "tags2" else []
        subprocess.run([nvcc, "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                        "-Wno-deprecated-gpu-targets", *extra, "-shared", "-o", str(out), str(src)], check=True,
                       capture_output=True)
    lib = ctypes.CDLL(str(out))
    lib.string_lookup_host_check.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                             ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]

    def run(words, ids_matrix, ld=None):
        rows = np.full((ids_matrix.shape[0],), -7, dtype=np.int64)
        ld = ids_matrix.strides[0] if ld is None else ld
        vec = lib.string_lookup_host_check(words.ctypes.data, words.shape[0], words.shape[1], ids_matrix.ctypes.data,
                                           ids_matrix.shape[0], ids_matrix.shape[1], ld, rows.ctypes.data)
        return rows, vec
    return run


def _aligned(a):
    """A 16-byte aligned C-contiguous copy of a uint8 matrix."""
    buf = np.zeros(a.size + 16, dtype=np.uint8)
    off = (-buf.ctypes.data) % 16
    out = buf[off:off + a.size].reshape(a.shape)
    out[...] = a
    return out


def test_pack_strings_layout(rs):
    m = rs.ops.pack_strings(["1", "10", "héllo", ""])
    assert m.dtype == np.uint8 and m.shape == (4, 16)
    assert bytes(m[2]).rstrip(b"\0") == "héllo".encode() and not m[3].any()
    assert rs.ops.pack_strings(np.array([b"ab", b"c"])).tolist() == rs.ops.pack_strings(["ab", "c"]).tolist()
    assert rs.ops.pack_strings(np.array(["ab", "c"])).tolist() == rs.ops.pack_strings(["ab", "c"]).tolist()
    assert rs.ops.pack_strings(["x" * 17]).shape == (1, 32) and rs.ops.pack_strings([]).shape == (0, 16)
    assert rs.ops.pack_strings(["abc"], width=48).shape == (1, 48)
    with pytest.raises(ValueError):
        rs.ops.pack_strings(["x" * 20], width=16)


def test_pack_str_list_equals_pack_strings_or_declines(rs):
    for _, vocab, ids in _cases():
        for seq in (list(vocab), list(ids), tuple(ids)):
            got = rs.ops.pack_str_list(seq)
            assert got is not None and got.dtype == np.uint8 and np.array_equal(got, rs.ops.pack_strings(seq))
    # a NUL inside a string, entries that are not str, missing values: the caller must keep the host dict
    assert rs.ops.pack_str_list(["ab", "a\0b"]) is None and rs.ops.pack_str_list(["a\0"]) is None
    assert rs.ops.pack_str_list(["a", 7]) is None
    assert np.array_equal(rs.ops.pack_str_list(["a", b"b"]), rs.ops.pack_strings(["a", b"b"]))   # bytes are taken as UTF-8, as lookup_host does
    assert rs.ops.pack_str_list(["a", None]) is None and rs.ops.pack_str_list(["a", "x" * 4097]) is None
    assert rs.ops.pack_str_list(["a", "x" * 4096]).shape == (2, 4096)


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c[0])
def test_lookup_routine_on_the_host_matches_the_reference_semantics(rs, host_check, case):
    _, vocab, ids = case
    want = _want(vocab, ids)
    words = rs.ops.pack_strings(vocab)
    # 16-byte rows, aligned: the vectorised path
    m = _aligned(rs.ops.pack_strings(ids))
    got, vec = host_check(words, m)
    assert vec == 1 and np.array_equal(got, want)
    # wider rows than the words (ids may be longer than every word) and a width that is not a multiple of 16: byte path
    w_odd = m.shape[1] + 5
    got, vec = host_check(words, np.ascontiguousarray(rs.ops.pack_strings(ids, width=w_odd)))
    assert vec == 0 and np.array_equal(got, want)
    # narrower rows than the pool (ids cut to their own longest string)
    if len(ids):
        own = max(len(s.encode()) for s in ids) or 1
        got, _ = host_check(words, np.ascontiguousarray(rs.ops.pack_strings(ids, width=own)))
        assert np.array_equal(got, want)
    # bytes after the first NUL are not part of the string (Arrow-style buffers may hold garbage there)
    dirty = m.copy()
    for r in range(dirty.shape[0]):
        n = int(np.argmax(dirty[r] == 0)) if (dirty[r] == 0).any() else dirty.shape[1]
        if n + 1 < dirty.shape[1]:
            dirty[r, n + 1:] = 0x41 + (r % 7)
    got, _ = host_check(words, _aligned(dirty))
    assert np.array_equal(got, want)


def test_lookup_routine_with_a_row_stride(rs, host_check):
    vocab, ids = ["a", "bb", "ccc"], ["bb", "zz", "ccc", "a", ""]
    wide = np.zeros((len(ids), 64), dtype=np.uint8)
    wide[:, :16] = rs.ops.pack_strings(ids)
    wide[:, 16:] = 0x5A                                              # the neighbouring columns of a wider buffer
    wide = _aligned(wide)
    got, vec = host_check(rs.ops.pack_strings(vocab), wide[:, :16], ld=64)
    assert vec == 1 and np.array_equal(got, _want(vocab, ids))


# ----------------------------------------------------------------------------------------------------------------------
# on the GPU, through the C ABI
# ----------------------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("case", _cases(), ids=lambda c: c[0])
def test_device_lookup_matches_the_reference_semantics(rs, case):
    import torch
    _, vocab, ids = case
    want = _want(vocab, ids)
    dv = rs.ops.DeviceVocabulary(vocab)
    assert len(dv) == len(vocab)
    m = rs.ops.pack_strings(ids)
    got = dv.lookup(torch.from_numpy(m).cuda())
    assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), want)
    odd = torch.from_numpy(rs.ops.pack_strings(ids, width=m.shape[1] + 5)).cuda()          # byte path
    assert np.array_equal(dv.lookup(odd).cpu().numpy(), want)
    wide = torch.full((len(ids), m.shape[1] + 48), 0x5A, dtype=torch.uint8, device="cuda")   # strided view of a wider buffer
    wide[:, :m.shape[1]] = torch.from_numpy(m).cuda()
    assert np.array_equal(dv.lookup(wide[:, :m.shape[1]]).cpu().numpy(), want)
    if len(ids):
        big = np.tile(m, (max(1, 70000 // len(ids)), 1))                                     # more ids than one grid pass
        got = dv.lookup(torch.from_numpy(big).cuda()).cpu().numpy()
        assert np.array_equal(got, np.tile(want, big.shape[0] // len(ids)))


@pytest.mark.gpu
def test_model_call_with_byte_ids_equals_the_string_call(rs):
    import torch
    cfg = rs.ModelConfig()
    n_users, n_items = 300, 200
    w = rs.synth.make_weights(cfg, n_users, n_items, seed=9)
    uv, iv = rs.synth.make_vocab(n_users), rs.synth.make_vocab(n_items)
    model = rs.MultiTaskModel(cfg, uv, iv, {})
    rs.load_weights(model, w)
    rng = np.random.default_rng(10)
    users = [uv[i] for i in rng.integers(0, n_users, 700)] + ["nobody", ""]
    items = [iv[i] for i in rng.integers(0, n_items, 700)] + ["9999", "7"]
    a = model({"user_id": users, "movie_id": items}, training=False)
    b = model({"user_id": np.array([s.encode() for s in users]), "movie_id": np.array([s.encode() for s in items])},
              training=False)
    c = model({"user_id": torch.from_numpy(rs.ops.pack_strings(users)).cuda(),
               "movie_id": torch.from_numpy(rs.ops.pack_strings(items))}, training=False)
    for key in a:
        assert torch.equal(a[key], b[key]) and torch.equal(a[key], c[key]), key
    enc = model.encoder
    assert np.array_equal(enc.user_lookup.lookup_device(np.array([s.encode() for s in users])).cpu().numpy(),
                          enc.user_lookup.lookup_host(users))
    with pytest.raises(RuntimeError):
        rs.ops.DeviceVocabulary(["a"]).lookup(torch.zeros((2, 16), dtype=torch.uint8, device="cuda"),
                                              out=torch.zeros((3,), dtype=torch.int64, device="cuda"))


@pytest.mark.gpu
def test_large_string_batches_against_large_vocabularies_take_the_device_path(rs, monkeypatch):
    """StringLookup.__call__ on a list of str: host dict by default, the device table for >= DEVICE_MIN_IDS ids against
    >= DEVICE_MIN_WORDS words (where the dict misses the CPU caches on every probe) -- same rows either way, and ids the
    byte layout cannot carry (a NUL inside) keep the whole call on the dict."""
    vocab = rs.synth.make_vocab(250_000)
    look = rs.StringLookup(vocab)
    rng = np.random.default_rng(3)
    ids = [vocab[i] for i in rng.integers(0, len(vocab), 9000)] + ["nobody", "", "héllo", "250001"]
    want = look.lookup_host(ids)
    small = look(ids[:100])
    assert look._device_vocab is None and np.array_equal(small.cpu().numpy(), want[:100])       # short list: host dict
    got = look(ids)
    assert look._device_vocab is not None and np.array_equal(got.cpu().numpy(), want)            # long list: device table
    odd = ids + ["1\x002"]
    assert np.array_equal(look(odd).cpu().numpy(), look.lookup_host(odd))                         # NUL inside: dict, 0 for that id
    monkeypatch.setenv("RS_DEVICE_LOOKUP", "0")
    assert np.array_equal(look(ids).cpu().numpy(), want)
    few_words = rs.StringLookup(vocab[:1000])
    few_words(ids)
    assert few_words._device_vocab is None                                                       # small vocabulary: host dict
