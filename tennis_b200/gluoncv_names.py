"""Structural parameter names of the GluonCV backbones, so that checkpoints written by the reference's
`model.save_parameters()` (train.py:497; models/README.md ids 0006 / 0042 / 0102) load into this package's models.

Gluon's `save_parameters` keys every tensor by its attribute path.  The reference's backbone is the zoo model's `.features`
HybridSequential (train.py:204), so inside `FrameModel` the keys read `backbone.<i>....` and inside `CNNRNN`
`td.model.<i>....`, where <i> is the child index in the zoo definition:

DenseNet-121 (gluoncv/model_zoo/densenet.py): 0 conv, 1 bn, [2 relu, 3 maxpool], 4/6/8/10 dense blocks, 5/7/9 transitions,
11 final bn; a dense layer is HybridConcurrent[Identity, Sequential(0 bn, 1 relu, 2 conv1x1, 3 bn, 4 relu, 5 conv3x3)] ->
`<blk>.<layer>.1.<0|2|3|5>.<param>`; a transition is Sequential(0 bn, 1 relu, 2 conv, 3 avgpool).
ResNet-18 v2 (resnetv2): 0 bn(scale=False, center=False), 1 conv, 2 bn, [3 relu, 4 maxpool], 5..8 stages of BasicBlockV2
(attributes bn1, conv1, bn2, conv2, downsample), 9 final bn.

[UPSTREAM, recalled; confidence M]: no GluonCV install or real checkpoint is available offline to confirm the indices; the one
name the reference itself documents (`td.model.4.0.1.0.gamma`, first BatchNorm of the first dense layer) agrees with this table.
"""

_DENSE_CFG = (6, 12, 24, 16)
_BN = ("gamma", "beta", "running_mean", "running_var")


def densenet121_name_map():
    """{GluonCV structural name (relative to `.features`): this package's name}, in inventory order."""
    m = {"0.weight": "conv0.weight"}
    m.update({"1.%s" % s: "bn0.%s" % s for s in _BN})
    idx = 4
    for b, nl in enumerate(_DENSE_CFG):
        for l in range(nl):
            src, dst = "%d.%d.1." % (idx, l), "block%d.layer%d." % (b + 1, l + 1)
            m.update({src + "0.%s" % s: dst + "bn1.%s" % s for s in _BN})
            m[src + "2.weight"] = dst + "conv1.weight"
            m.update({src + "3.%s" % s: dst + "bn2.%s" % s for s in _BN})
            m[src + "5.weight"] = dst + "conv2.weight"
        idx += 1
        if b < 3:
            m.update({"%d.0.%s" % (idx, s): "trans%d.bn.%s" % (b + 1, s) for s in _BN})
            m["%d.2.weight" % idx] = "trans%d.conv.weight" % (b + 1)
            idx += 1
    m.update({"%d.%s" % (idx, s): "bn5.%s" % s for s in _BN})
    return m


def resnet18_v2_name_map():
    m = {"0.%s" % s: "bn_data.%s" % s for s in _BN}
    m["1.weight"] = "conv0.weight"
    m.update({"2.%s" % s: "bn0.%s" % s for s in _BN})
    cin = 64
    for s_, c in enumerate((64, 128, 256, 512)):
        for b in range(2):
            src, dst = "%d.%d." % (5 + s_, b), "stage%d.block%d." % (s_ + 1, b + 1)
            m.update({src + "bn1.%s" % s: dst + "bn1.%s" % s for s in _BN})
            m[src + "conv1.weight"] = dst + "conv1.weight"
            m.update({src + "bn2.%s" % s: dst + "bn2.%s" % s for s in _BN})
            m[src + "conv2.weight"] = dst + "conv2.weight"
            if b == 0 and cin != c:
                m[src + "downsample.weight"] = dst + "downsample.weight"
            cin = c
    m.update({"9.%s" % s: "bn_final.%s" % s for s in _BN})
    return m


NAME_MAPS = {"densenet121": densenet121_name_map, "resnet18_v2": resnet18_v2_name_map}


def translate_checkpoint_keys(loaded, params):
    """Rename GluonCV-structural keys in `loaded` (dict name -> array) to this package's names wherever a `Features` block of
    `params`' model sits: `<prefix><gluoncv name>` -> `<prefix><our name>` for every prefix under which our inventory appears
    (e.g. 'backbone.', 'td.model.').  Keys that already match are left alone."""
    prefixes = {}
    for key in params:
        for arch, first in (("densenet121", "block1.layer1.bn1.gamma"), ("resnet18_v2", "stage1.block1.bn1.gamma")):
            if key.endswith(first):
                prefixes[key[:-len(first)]] = arch
    out = {}
    for key, arr in loaded.items():
        new = key
        if key not in params and "._proj_query." in key:
            # gluonnlp's attention cell keeps its query projection in a private attribute (confidence M)
            new = key.replace("._proj_query.", ".proj_query.")
        if new not in params:
            for prefix, arch in prefixes.items():
                if new.startswith(prefix):
                    mapped = NAME_MAPS[arch]().get(new[len(prefix):])
                    if mapped is not None:
                        new = prefix + mapped
                        break
        out[new] = arr
    return out
