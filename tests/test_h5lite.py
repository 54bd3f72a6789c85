"""Keras .h5 weight layout through the hand-written HDF5 subset: round trip + structural checks against the
published file format (no foreign HDF5 implementation exists in this environment — see h5lite.py)."""
import struct

import numpy as np
import pytest

from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, h5lite, spec_from_config, weights


@pytest.mark.parametrize("name", ["h36m_351", "h36m_81"])
def test_roundtrip_and_layout(tmp_path, name):
    spec = spec_from_config(UpliftUpsampleConfig.preset(name))
    w = weights.init_weights(spec, 5, perturb=True)
    p = str(tmp_path / "w.h5")
    h5lite.save_keras_weights(p, spec, w)
    back = h5lite.load_keras_weights(p, spec)
    assert set(back) == set(w) and all(np.array_equal(back[k], w[k]) for k in w)
    raw = open(p, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    (eof,) = struct.unpack_from("<Q", raw, 40)
    assert eof == len(raw)
    f = h5lite.H5File(p)
    layers = [x.decode() for x in f.attrs["layer_names"]]
    assert layers == list(weights.inventory(spec))
    assert f.attrs["backend"] == b"tensorflow" and f.attrs["keras_version"] == b"2.4.0"
    assert len(f["token_dropout"].attrs["weight_names"]) == 0
    g = f["spatial_block_1"]
    names = [x.decode() for x in g.attrs["weight_names"]]
    assert names[0] == "spatial_block_1/layer_normalization/gamma:0" and len(names) == 16
    # nested groups: /spatial_block_1/spatial_block_1/mha/dense/kernel:0
    assert g["spatial_block_1/mha/dense/kernel:0"].read().shape == (32, 32)
    assert f["strided_temporal_block_1"][names[14].replace("spatial_block_1", "strided_temporal_block_1")
                                          .replace("mlp/dense_5", "strided_mlp/conv1d_1")].read().shape == (3, 768, 384)
    # every symbol node is sorted bytewise and every object header is 8-byte aligned
    links = f.links()
    assert list(links) == sorted(links, key=lambda s: s.encode()) and all(a % 8 == 0 for a in links.values())


def test_model_weights_wrapper_and_errors(tmp_path):
    spec = spec_from_config(UpliftUpsampleConfig.preset("h36m_81"))
    w = weights.init_weights(spec, 2)
    inv = weights.inventory(spec)
    root = h5lite._Node()
    mw = root.ensure_group(["model_weights"])                      # whole-model file: weights under /model_weights
    mw.attrs = [("layer_names", list(inv)), ("backend", b"tensorflow"), ("keras_version", b"2.4.0")]
    for gname, tensors in inv.items():
        grp = mw.ensure_group([gname])
        grp.attrs = [("weight_names", [n for n, _, _ in tensors])]
        for i, (wn, _, _) in enumerate(tensors):
            parts = wn.split("/")
            grp.ensure_group(parts[:-1]).children[parts[-1]] = h5lite._Node(w[(gname, i)])
    p = str(tmp_path / "model.h5")
    h5lite.write_h5(p, root)
    back = h5lite.load_keras_weights(p, spec)
    assert all(np.array_equal(back[k], w[k]) for k in w)
    # shape / count mismatches raise, like weight_io.py:185-232
    spec351 = spec_from_config(UpliftUpsampleConfig.preset("h36m_351"))
    with pytest.raises(ValueError, match="shape"):
        h5lite.load_keras_weights(p, spec351)
    # ... unless skip_mismatch: the mismatching tensors are skipped and reported, the matching ones load (weight_io.py:186, :220)
    rep = {}
    part = h5lite.load_keras_weights(p, spec351, skip_mismatch=True, verbose=False, report=rep)
    assert rep["skipped"] and 0 < len(part) < len(w) and all(np.array_equal(part[k], w[k]) for k in part)
    bad = tmp_path / "bad.h5"
    bad.write_bytes(b"not hdf5 at all")
    with pytest.raises(ValueError, match="signature"):
        h5lite.H5File(str(bad))


def test_groups_with_many_links_use_several_symbol_nodes(tmp_path):
    root = h5lite._Node()
    for i in range(37):
        root.children[f"d{i:02d}"] = h5lite._Node(np.full((2, 3), i, np.float32))
    root.attrs = [("note", b"x")]
    p = str(tmp_path / "many.h5")
    h5lite.write_h5(p, root)
    f = h5lite.H5File(p)
    assert len(f.links()) == 37 and f["d36"].read()[0, 0] == 36 and f.attrs["note"] == b"x"


def test_layers_missing_from_the_file_are_reported_not_fatal(tmp_path, capsys):
    """By-name loading (weight_io.py:241-251): a model layer the file does not hold keeps its values and is listed under "not
    assigned any weights"; a file layer the model does not have is listed as "not consumed"."""
    spec = spec_from_config(UpliftUpsampleConfig.preset("h36m_81"))
    w = weights.init_weights(spec, 3)
    p = str(tmp_path / "full.h5")
    h5lite.save_keras_weights(p, spec, w)
    f = h5lite.H5File(p)
    layers = [h5lite._decode(n) for n in np.atleast_1d(f.attrs["layer_names"])]
    drop = "temporal_fc"
    assert drop in layers
    # rebuild the file without one layer and with one foreign layer
    root = h5lite._Node()
    names = [n for n in layers if n != drop] + ["some_other_head"]
    root.attrs = [("layer_names", names), ("backend", b"tensorflow"), ("keras_version", b"2.4.0")]
    inv = weights.inventory(spec)
    for gname in names:
        foreign = gname == "some_other_head"
        tensors = [("some_other_head/kernel:0", (3, 3), None)] if foreign else inv[gname]
        grp = root.ensure_group([gname])
        grp.attrs = [("weight_names", [t[0] for t in tensors])]
        for i, (wn, shape, _) in enumerate(tensors):
            parts = wn.split("/")
            arr = np.zeros(shape, np.float32) if foreign else w[(gname, i)]
            grp.ensure_group(parts[:-1]).children[parts[-1]] = h5lite._Node(arr)
    q = str(tmp_path / "partial.h5")
    h5lite.write_h5(q, root)
    rep = {}
    back = h5lite.load_keras_weights(q, spec, report=rep)
    out = capsys.readouterr().out
    assert rep["unassigned_layers"] == [drop] and rep["unconsumed_layers"] == ["some_other_head"]
    assert "not assigned any weights" in out and "- " + drop in out and "not consumed" in out
    assert all(k[0] != drop for k in back) and all(np.array_equal(back[k], w[k]) for k in back)
    assert len(back) == len(w) - len(inv[drop])
