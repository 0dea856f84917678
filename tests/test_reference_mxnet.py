"""Pins the CPU oracle against the UNTOUCHED reference classes wherever MXNet exists.

The oracle (oracle/vision.py, oracle/captioning.py) is a torch restatement of the Gluon graphs; this image has no mxnet / gluoncv /
gluonnlp wheel and no network, so the pin cannot run here and the test SKIPS.  On a machine where `import mxnet, gluoncv` works and
the reference checkout is reachable (TENNIS_REFERENCE_DIR, ./baseline/_ref or /root/reference), it builds the reference's own
`FrameModel` / `CNNRNN` (models/vision/definitions.py:10-33, 75-110) on the GluonCV DenseNet-121, loads the seeded synthetic
parameters through the GluonCV structural names (tennis_b200/gluoncv_names.py), runs the reference forward on the CPU context
(evaluate.py:85 `mx.cpu()`), and asserts the oracle reproduces its features and logits to fp32 rounding.  CPU-only, no GPU marker."""
import os
import sys

import numpy as np
import pytest

mx = pytest.importorskip("mxnet", reason="MXNet is not installable offline: the oracle stays 'parity unpinned' (DESIGN.md section 2)")
pytest.importorskip("gluoncv", reason="GluonCV model zoo not available")


def _reference_dir():
    for d in (os.environ.get("TENNIS_REFERENCE_DIR"), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                    "baseline", "_ref"), "/root/reference"):
        if d and os.path.exists(os.path.join(d, "models", "vision", "definitions.py")):
            return d
    return None


@pytest.mark.skipif(_reference_dir() is None, reason="reference checkout not reachable")
def test_reference_cnnrnn_on_mxnet_cpu_matches_the_oracle():
    import torch
    from gluoncv.model_zoo import get_model
    from oracle import vision as O
    from tennis_b200 import synthetic as S
    from tennis_b200.gluoncv_names import densenet121_name_map
    sys.path.insert(0, _reference_dir())
    from models.vision.definitions import CNNRNN, FrameModel  # the reference's own classes, unmodified
    ctx = mx.cpu()
    backbone = get_model("DenseNet121", pretrained=False, ctx=ctx).features  # train.py:204
    model = CNNRNN(FrameModel(backbone, 11), 11, type="gru", hidden_size=128)  # train.py:236
    model.initialize(ctx=ctx)
    T, size = 4, 224
    u8, clips = S.structured_clips(2, T, size, seed=300)
    x = mx.nd.array(clips.numpy(), ctx=ctx)
    model(x)  # resolve deferred shapes
    p = S.synthetic_params("densenet121", seed=1234)
    rp = S.synthetic_rnn_params("gru", 1024, 128, seed=4321)
    params = model.collect_params()
    feat_prefix = [k for k in params.keys() if k.endswith("conv0_weight") or k.endswith("0.weight")]
    # Gluon's structural names (save_parameters / _collect_params_with_prefix): td.model.<gluoncv index path>
    structural = model._collect_params_with_prefix()
    for gname, ours in densenet121_name_map().items():
        structural["td.model." + gname].set_data(mx.nd.array(p[ours].numpy()))
    for k, v in rp.items():
        structural["rnn." + k].set_data(mx.nd.array(v.numpy()))
    g = torch.Generator().manual_seed(77)
    cw = (torch.rand(11, 256, generator=g) * 2 - 1) * 0.07
    structural["classes.weight"].set_data(mx.nd.array(cw.numpy()))
    structural["classes.bias"].set_data(mx.nd.zeros((11,)))
    assert feat_prefix is not None
    ref_logits = model(x).asnumpy()
    ref_feats = backbone(x.reshape((-1, 3, size, size))).asnumpy().reshape(2 * T, -1)
    with torch.no_grad():
        of = O.FEATURES["densenet121"](clips.reshape(-1, 3, size, size), p).numpy()
        ol = O.cnnrnn(clips, lambda z: O.FEATURES["densenet121"](z, p), rp, "gru", 128, cw, torch.zeros(11)).numpy()
    assert np.abs(of - ref_feats).max() < 1e-3 * max(1.0, np.abs(ref_feats).max())
    assert np.abs(ol - ref_logits).max() < 1e-4
    assert (ol.argmax(1) == ref_logits.argmax(1)).all()
