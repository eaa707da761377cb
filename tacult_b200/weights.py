"""state_dict -> weight blob for tc_load_weights.

Key names and shapes are the reference's own (`_UtacNNet`, /root/reference/src/tacult/network.py:44-79;
listed in SURVEY.md §11.2), so `nnet.nnet.state_dict()` of a reference `NNetWrapper` can be passed in
unchanged (torch tensors or numpy arrays).  BatchNorm folding happens inside the library.
"""
import numpy as np

from ._lib import NetDesc


def _np(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.float32)


def _strip(state_dict):
    # torch.compile'd modules prefix keys with "_orig_mod."
    out = {}
    for k, v in state_dict.items():
        out[k[len("_orig_mod."):] if k.startswith("_orig_mod.") else k] = v
    return out


ENGINE_PLANES = ("side-to-move stones", "opponent stones", "legal cells", "ones")


def state_dict_to_blob(state_dict, plane_order=None):
    """plane_order[i] = which ENGINE plane (ENGINE_PLANES) the network's input plane i was trained on.  The engine always
    builds its four planes in ENGINE_PLANES order (the reference does not pin `get_obs()`), so a
    checkpoint trained on another ordering is served by permuting conv_in's input channels here — the forward result is
    the same as feeding the planes in the checkpoint's own order.  None = identity."""
    sd = dict(_strip(state_dict))
    if plane_order is not None:
        order = [int(x) for x in plane_order]
        w = _np(sd["conv_in.weight"])
        if sorted(order) != list(range(w.shape[1])):
            raise ValueError(f"plane_order must be a permutation of 0..{w.shape[1] - 1}, got {order}")
        permuted = np.empty_like(w)
        for net_plane, engine_plane in enumerate(order):
            permuted[:, engine_plane] = w[:, net_plane]
        sd["conv_in.weight"] = permuted
    conv_in = _np(sd["conv_in.weight"])
    channels, in_planes = conv_in.shape[0], conv_in.shape[1]
    blocks = 0
    while f"resnet.{blocks}.conv_block1.0.weight" in sd:
        blocks += 1
    emb = _np(sd["fc1.weight"]).shape[0]
    if _np(sd["fc1.weight"]).shape[1] != 81 * channels:
        raise ValueError("fc1.weight does not match 81 * channels inputs")
    parts = []

    def layer(prefix):
        parts.append(_np(sd[prefix + ".weight"]).ravel())
        parts.append(_np(sd[prefix + ".bias"]).ravel())

    def bn(prefix):
        for name in ("weight", "bias", "running_mean", "running_var"):
            parts.append(_np(sd[f"{prefix}.{name}"]).ravel())

    layer("conv_in")
    bn("bn_in")
    for k in range(blocks):
        layer(f"resnet.{k}.conv_block1.0")
        bn(f"resnet.{k}.conv_block1.1")
        layer(f"resnet.{k}.conv_block2.0")
        bn(f"resnet.{k}.conv_block2.1")
    layer("fc1")
    bn("fc_bn1")
    layer("pi")
    layer("v")
    blob = np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)
    return NetDesc(int(channels), int(blocks), int(emb), int(in_planes)), blob


def random_state_dict(seed, channels=32, num_residual_blocks=3, embedding_size=256, in_planes=4):
    """Random-init weights with the reference's state_dict layout (synthetic benchmark input;
    He-scaled convolutions, non-trivial BatchNorm statistics so that folding is exercised)."""
    rng = np.random.RandomState(seed)
    sd = {}

    def conv(prefix, cin, cout):
        sd[prefix + ".weight"] = (rng.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / (cin * 9))).astype(np.float32)
        sd[prefix + ".bias"] = (rng.standard_normal(cout) * 0.05).astype(np.float32)

    def bn(prefix, n):
        sd[prefix + ".weight"] = rng.uniform(0.6, 1.4, n).astype(np.float32)
        sd[prefix + ".bias"] = (rng.standard_normal(n) * 0.1).astype(np.float32)
        sd[prefix + ".running_mean"] = (rng.standard_normal(n) * 0.1).astype(np.float32)
        sd[prefix + ".running_var"] = rng.uniform(0.5, 1.5, n).astype(np.float32)

    def linear(prefix, fin, fout, gain):
        sd[prefix + ".weight"] = (rng.standard_normal((fout, fin)) * gain / np.sqrt(fin)).astype(np.float32)
        sd[prefix + ".bias"] = (rng.standard_normal(fout) * 0.05).astype(np.float32)

    conv("conv_in", in_planes, channels)
    bn("bn_in", channels)
    for k in range(num_residual_blocks):
        conv(f"resnet.{k}.conv_block1.0", channels, channels)
        bn(f"resnet.{k}.conv_block1.1", channels)
        conv(f"resnet.{k}.conv_block2.0", channels, channels)
        bn(f"resnet.{k}.conv_block2.1", channels)
    linear("fc1", 81 * channels, embedding_size, np.sqrt(2.0))
    bn("fc_bn1", embedding_size)
    linear("pi", embedding_size, 81, 2.0)
    linear("v", embedding_size, 1, 1.0)
    return sd


def flops_per_eval(channels, blocks, emb, in_planes=4):
    """Algorithmic FLOPs of one leaf evaluation."""
    return 2 * 81 * 9 * channels * (in_planes + 2 * blocks * channels) + 2 * 81 * channels * emb + 2 * emb * 82
