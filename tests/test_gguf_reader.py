"""The GGUF container reader (utils/gguf_file.py) against files written by gguf-py (test
infrastructure only: the reader itself is written from the format description and imports nothing from gguf-py)."""
import os
import struct

import numpy as np
import pytest

gguf = pytest.importorskip("gguf")

from oracle import ggq_oracle as orc                      # noqa: E402
from utils.gguf_file import GGUFError, GGUFFile           # noqa: E402

QT = gguf.GGMLQuantizationType
SHAPES = {"q8_0": (48, 256, QT.Q8_0), "q4_k": (32, 512, QT.Q4_K), "q6_k": (24, 768, QT.Q6_K)}


@pytest.fixture(scope="module")
def model_file(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("gguf") / "tiny.gguf")
    w = gguf.GGUFWriter(path, "llama")
    w.add_uint32("llama.block_count", 2)
    w.add_float32("llama.attention.layer_norm_rms_epsilon", 1e-5)
    w.add_string("general.name", "tiny-test")
    w.add_bool("tokenizer.ggml.add_bos_token", True)
    w.add_array("tokenizer.ggml.tokens", ["<s>", "</s>", "héllo"])
    w.add_array("tokenizer.ggml.scores", [0.0, -1.5, 2.25])
    packed = {}
    rng = np.random.default_rng(5)
    norm = rng.standard_normal(256).astype(np.float32)
    emb = rng.standard_normal((8, 64)).astype(np.float16)
    w.add_tensor("output_norm.weight", norm)
    w.add_tensor("token_embd.weight", emb)
    for fmt, (O, K, qt) in SHAPES.items():
        blocks = orc.random_blocks(fmt, O, K, seed=11)
        rowb = orc.packed_nbytes(fmt, 1, K)
        packed[fmt] = blocks
        w.add_tensor(f"blk.0.{fmt}.weight", blocks.view(np.uint8).reshape(O, rowb), raw_dtype=qt)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return path, packed, norm, emb


def test_metadata_and_tensor_table(model_file):
    path, packed, norm, emb = model_file
    with GGUFFile(path) as g:
        assert g.version == 3 and g.alignment == 32 and g.data_start % 32 == 0
        md = g.metadata
        assert md["general.architecture"] == "llama" and md["general.name"] == "tiny-test"
        assert md["llama.block_count"] == 2 and md["tokenizer.ggml.add_bos_token"] is True
        assert abs(md["llama.attention.layer_norm_rms_epsilon"] - 1e-5) < 1e-12
        assert md["tokenizer.ggml.tokens"] == ["<s>", "</s>", "héllo"]
        assert md["tokenizer.ggml.scores"] == [0.0, -1.5, 2.25]
        assert set(g.tensors) == {"output_norm.weight", "token_embd.weight"} | {f"blk.0.{f}.weight" for f in SHAPES}
        assert np.array_equal(g.tensor_numpy("output_norm.weight"), norm)
        assert g.tensors["token_embd.weight"].shape == (8, 64) and np.array_equal(g.tensor_numpy("token_embd.weight"), emb)
        for fmt, (O, K, qt) in SHAPES.items():
            t = g.tensors[f"blk.0.{fmt}.weight"]
            assert t.shape == (O, K) and t.mmq_format == fmt and t.type_name == qt.name
            assert t.nbytes == orc.packed_nbytes(fmt, O, K) and t.offset % 32 == 0
            # the bytes are exactly the flat stream the reference packers emit / the mmq ops consume
            assert np.array_equal(g.tensor_bytes(t.name), packed[fmt].view(np.uint8).ravel())


def test_agrees_with_gguf_py_reader_and_dequantizer(model_file):
    path, packed, _, _ = model_file
    theirs = {t.name: t for t in gguf.GGUFReader(path).tensors}
    with GGUFFile(path) as g:
        for name, t in g.tensors.items():
            assert t.offset == theirs[name].data_offset and t.nbytes == theirs[name].n_bytes
            assert t.ggml_type == int(theirs[name].tensor_type)
        for fmt, (O, K, qt) in SHAPES.items():
            mine = orc.dequantize(fmt, g.tensor_bytes(f"blk.0.{fmt}.weight").view(np.int8), (O, K))
            ref = gguf.quants.dequantize(np.asarray(theirs[f"blk.0.{fmt}.weight"].data), qt).reshape(O, K)
            assert np.array_equal(np.asarray(mine, dtype=np.float32).astype(np.float16), ref.astype(np.float16))


def test_type_table_matches_gguf_py():
    from utils.gguf_file import GGML_TYPES
    for tid, (name, qk, blk) in GGML_TYPES.items():
        qt = QT(tid)
        assert qt.name == name and gguf.GGML_QUANT_SIZES[qt] == (qk, blk)


def test_rejects_garbage(tmp_path):
    bad = tmp_path / "bad.gguf"
    bad.write_bytes(b"GGML" + b"\0" * 64)
    with pytest.raises(GGUFError):
        GGUFFile(str(bad))
    bad.write_bytes(b"GGUF" + struct.pack("<IQQ", 1, 0, 0))            # version 1
    with pytest.raises(GGUFError):
        GGUFFile(str(bad))
    bad.write_bytes(b"GGUF" + struct.pack("<IQQ", 3, 1, 0) + struct.pack("<Q", 4) + b"name" + struct.pack("<IQ", 2, 256))
    with pytest.raises(GGUFError):                                     # tensor info cut short
        GGUFFile(str(bad))
    # a Q4_K tensor whose rows are not whole blocks
    hdr = b"GGUF" + struct.pack("<IQQ", 3, 1, 0) + struct.pack("<Q", 1) + b"w" + struct.pack("<IQQIQ", 2, 100, 4, 12, 0)
    bad.write_bytes(hdr + b"\0" * 4096)
    with pytest.raises(GGUFError):
        GGUFFile(str(bad))


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", list(SHAPES))
def test_linear_from_file_matches_oracle(model_file, fmt):
    import torch
    path, packed, _, _ = model_file
    O, K, _ = SHAPES[fmt]
    with GGUFFile(path) as g:
        lin = g.linear(f"blk.0.{fmt}.weight", "cuda:0")
        assert (lin.fmt, lin.O, lin.K) == (fmt, O, K)
        X = np.random.default_rng(3).standard_normal((5, K)).astype(np.float16)
        C = lin(torch.from_numpy(X).cuda()).cpu().numpy()
    ref = orc.ref32(fmt, packed[fmt], X, O, 5, K)
    mx, fro = orc.tier1_errors(ref, C.astype(np.float32))
    assert mx <= orc.TIER1_MAX and fro <= orc.TIER1_FRO, (mx, fro)
