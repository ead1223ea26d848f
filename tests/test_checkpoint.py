"""Checkpoint reader / writer (mreserve/checkpoint.py:29-119; flax msgpack wire format restated in
merlot_reserve_b200/checkpoint.py).  CPU only, except the last test which runs a restored tree on the GPU."""
import os

import ml_dtypes
import msgpack
import numpy as np
import pytest

from merlot_reserve_b200 import checkpoint as ck
from merlot_reserve_b200.synthetic import synthetic_clip, synthetic_params, tiny_config


def _state(seed=0):
    rng = np.random.default_rng(seed)
    params = {"enc": {"kernel": rng.standard_normal((4, 6)).astype(np.float32),
                      "bias": rng.standard_normal(6).astype(np.float32)},
              "contrastive_scales": np.ones(3, np.float32)}
    mom = ck.tree_map(lambda x: (x * 0.01).astype(ml_dtypes.bfloat16), params)
    return {"step": 7500, "params": params, "opt_state": {"0": {"count": np.int32(7500), "mu": mom, "nu": mom}}}


def test_wire_format_known_answer():
    """Byte-level layout of one leaf: fixmap{ 'a': ext8(type 1, msgpack([shape], dtype name, raw bytes)) }."""
    arr = np.array([1.0, -2.0], np.float16)
    payload = msgpack.packb(([2], "float16", arr.tobytes()), use_bin_type=True)
    expect = b"\x81\xa1a" + b"\xc7" + bytes([len(payload)]) + b"\x01" + payload
    assert ck.msgpack_serialize({"a": arr}) == expect
    # payload itself: array(3) | array(1)[2] | str 'float16' | bin8 4 bytes (little-endian halves 0x3c00, 0xc000)
    assert payload == b"\x93\x91\x02\xa7float16\xc4\x04\x00\x3c\x00\xc0"
    back = ck.msgpack_restore(expect)
    assert back["a"].dtype == np.float16 and np.array_equal(back["a"], arr)


def test_scalars_none_and_bfloat16_round_trip():
    tree = {"step": 3, "none": None, "s": np.float32(1.5), "i": np.int32(-4), "c": 1 + 2j,
            "b": np.arange(6, dtype=np.float32).astype(ml_dtypes.bfloat16).reshape(2, 3)}
    back = ck.msgpack_restore(ck.msgpack_serialize(tree))
    assert back["step"] == 3 and back["none"] is None and back["c"] == 1 + 2j
    assert back["s"] == np.float32(1.5) and back["s"].dtype == np.float32 and back["i"] == -4
    assert back["b"].dtype == ml_dtypes.bfloat16 and back["b"].shape == (2, 3)
    assert np.array_equal(back["b"].astype(np.float32), np.arange(6, dtype=np.float32).reshape(2, 3))


def test_chunked_arrays_reassemble():
    a = np.arange(1000, dtype=np.float32).reshape(10, 100)
    enc = ck.msgpack_serialize({"w": a, "small": np.zeros(3, np.float32)}, max_chunk_bytes=1024)
    raw = msgpack.unpackb(enc, ext_hook=ck._ext_unpack, raw=False)
    # flax.serialization._chunk writes the shape through _tuple_to_dict: {'0': 10, '1': 100}
    assert raw["w"]["__msgpack_chunked_array__"] is True and raw["w"]["shape"] == {"0": 10, "1": 100}
    assert len(raw["w"]["chunks"]) == 4 and raw["w"]["chunks"]["0"].shape == (256,)
    assert isinstance(raw["small"], np.ndarray)
    assert np.array_equal(ck.msgpack_restore(enc)["w"], a)
    # the list form written by earlier versions of this file still reads
    raw["w"]["shape"] = [10, 100]
    assert np.array_equal(ck._unchunk_leaves(raw)["w"], a)


def test_save_load_compresses_fp32_to_fp16_and_keeps_bf16(tmp_path):
    st = _state()
    fn = ck.save_checkpoint(st, str(tmp_path))
    assert os.path.basename(fn) == "ckpt_7500"
    on_disk = ck.msgpack_restore(open(fn, "rb").read())
    assert on_disk["params"]["enc"]["kernel"].dtype == np.float16            # checkpoint.py:29-36
    assert on_disk["opt_state"]["0"]["mu"]["enc"]["kernel"].dtype == ml_dtypes.bfloat16
    for target in (None, _state(seed=5)):
        got = ck.load_checkpoint(str(tmp_path), state=target)
        assert got["step"] == 7500
        k = got["params"]["enc"]["kernel"]
        assert k.dtype == np.float32                                         # checkpoint.py:39-40
        assert np.array_equal(k, st["params"]["enc"]["kernel"].astype(np.float16).astype(np.float32))
        assert got["opt_state"]["0"]["nu"]["enc"]["bias"].dtype == ml_dtypes.bfloat16
    # file path instead of directory: what PretrainedMerlotReserve.from_pretrained passes (modeling.py:991)
    assert np.array_equal(ck.load_checkpoint(fn)["params"]["enc"]["bias"], got["params"]["enc"]["bias"])
    bf = ck.load_checkpoint(fn, use_bfloat16_weights=True)["params"]["enc"]["kernel"]
    assert bf.dtype == ml_dtypes.bfloat16


def test_step_selection_keep_and_overwrite(tmp_path):
    for s in (100, 200, 300):
        ck.save_checkpoint(dict(_state(), step=s), str(tmp_path), keep=2)
    assert sorted(os.listdir(tmp_path)) == ["ckpt_200", "ckpt_300"]
    assert ck.load_checkpoint(str(tmp_path))["step"] == 300
    assert ck.load_checkpoint(str(tmp_path), step=200)["step"] == 200
    with pytest.raises(ValueError):
        ck.load_checkpoint(str(tmp_path), step=100)
    with pytest.raises(ValueError):
        ck.save_checkpoint(dict(_state(), step=250), str(tmp_path), overwrite=False)
    ck.save_checkpoint(dict(_state(), step=250), str(tmp_path))               # overwrite drops the newer file
    assert sorted(os.listdir(tmp_path)) == ["ckpt_200", "ckpt_250"]
    empty = tmp_path / "nothing"
    empty.mkdir()
    st = _state()
    assert ck.load_checkpoint(str(empty), state=st)["step"] == 7500           # nothing found: target comes back
    with pytest.raises(FileNotFoundError):
        ck.load_checkpoint(str(empty))


def test_structure_mismatch_and_no_optimizer(tmp_path):
    ck.save_checkpoint(_state(), str(tmp_path), no_optimizer=True)
    got = ck.load_checkpoint(str(tmp_path))
    assert got["opt_state"] is None
    bad = _state()
    bad["params"]["extra"] = np.zeros(2, np.float32)
    with pytest.raises(ValueError):
        ck.load_checkpoint(str(tmp_path), state=bad)


def test_tree_helpers():
    st = _state()
    flat = ck.flatten_dict(st["params"])
    assert sorted(flat) == ["contrastive_scales", "enc/bias", "enc/kernel"]
    seen = []
    out = ck.tree_map_nested_keys(lambda k, v: seen.append(k) or v * 2, st["params"])
    assert sorted(seen) == sorted(flat) and np.array_equal(out["enc"]["bias"], st["params"]["enc"]["bias"] * 2)
    assert ck.log_param_shapes(st["params"]) == 4 * 6 + 6 + 3
    assert ck.bf16_to_f32(ck.f32_to_bf16(st["params"]))["enc"]["kernel"].dtype == np.float32


def test_full_parameter_tree_round_trip(tmp_path):
    cfg = tiny_config()
    params = synthetic_params(cfg, seed=0)
    ck.save_checkpoint({"step": 1, "params": params, "opt_state": None}, str(tmp_path))
    got = ck.load_checkpoint(str(tmp_path))["params"]
    a, b = ck.flatten_dict(params), ck.flatten_dict(got)
    assert sorted(a) == sorted(b)
    for k in a:
        assert b[k].shape == np.asarray(a[k]).shape and b[k].dtype == np.float32, k
        np.testing.assert_allclose(b[k], a[k], rtol=1e-3, atol=1e-7)          # one fp16 rounding


@pytest.mark.gpu
def test_restored_checkpoint_runs_on_the_b200_path(tmp_path):
    """save -> load -> PretrainedMerlotReserve.from_params: same embeddings as the fp16-rounded tree."""
    from merlot_reserve_b200 import PretrainedMerlotReserve
    cfg = tiny_config()
    params = synthetic_params(cfg, seed=0, weight_scale=4.0)
    ck.save_checkpoint({"step": 9, "params": params, "opt_state": None}, str(tmp_path))
    restored = ck.load_checkpoint(str(tmp_path))["params"]
    rounded = ck.tree_map(lambda x: np.asarray(x).astype(np.float16).astype(np.float32), params)
    clip = synthetic_clip(cfg, nseg=4, seed=3)
    got = PretrainedMerlotReserve.from_params(cfg, restored).embed_video(**clip)
    ref = PretrainedMerlotReserve.from_params(cfg, rounded).embed_video(**clip)
    assert np.array_equal(np.asarray(got), np.asarray(ref))


def test_tokenizer_hook_loads_the_reference_vocabulary():
    """get_encoder() + the padding protocol `get_label_space` relies on (modeling.py:1024-1032).  The vocabulary is a
    data file of the reference: the test runs where that tree is mounted and is skipped elsewhere (GPU box)."""
    from merlot_reserve_b200 import lowercase_encoder as le
    fn = "/root/reference/mreserve/lowercase_encoder.json"
    if not os.path.exists(fn):
        with pytest.raises(FileNotFoundError):
            le.get_encoder("/nonexistent/lowercase_encoder.json")
        pytest.skip("reference vocabulary not available")
    enc = le.get_encoder(fn)
    for i, tok in enumerate(le.SPECIAL_TOKENS[:7]):
        assert enc.token_to_id(tok) == i                      # lowercase_encoder.py:9-18
    enc.enable_padding(pad_token="<|PAD|>", length=15)
    ids = np.array([x.ids[:15] for x in enc.encode_batch(["cooking pasta", "playing the guitar on stage"])])
    enc.no_padding()
    assert ids.shape == (2, 15) and (ids[0] != le.PADDING).sum() < 15 and ids[0, -1] == le.PADDING
    assert (ids >= 0).all() and ids.max() < 32768
