"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU fp32 restatement of the reference's dual-head ResNet forward in eval mode
(/root/reference/src/tacult/network.py:44-102, ResidualBlock 22-41) and of the
`NNetWrapper.predict` boundary (/root/reference/src/tacult/base/nn_wrapper.py:125-136).
Weights use the reference's own state_dict key names, so a
state_dict taken from a real `_UtacNNet` can be fed in unchanged.

Pinned against the live `_UtacNNet` by tests/golden/make_golden.py (fixtures
net_*.npz).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm default, network.py:27,32,61,72


def make_state_dict(seed, channels=32, num_residual_blocks=3, embedding_size=256, in_planes=4):
    """Seeded numpy weights (platform independent), with non-trivial BN statistics."""
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
        sd[prefix + ".num_batches_tracked"] = np.array(1, dtype=np.int64)

    def linear(prefix, fin, fout, gain=1.0):
        sd[prefix + ".weight"] = (rng.standard_normal((fout, fin)) * gain / np.sqrt(fin)).astype(np.float32)
        sd[prefix + ".bias"] = (rng.standard_normal(fout) * 0.05).astype(np.float32)

    conv("conv_in", in_planes, channels)
    bn("bn_in", channels)
    for k in range(num_residual_blocks):
        conv(f"resnet.{k}.conv_block1.0", channels, channels)
        bn(f"resnet.{k}.conv_block1.1", channels)
        conv(f"resnet.{k}.conv_block2.0", channels, channels)
        bn(f"resnet.{k}.conv_block2.1", channels)
    linear("fc1", 81 * channels, embedding_size, gain=np.sqrt(2.0))
    bn("fc_bn1", embedding_size)
    linear("pi", embedding_size, 81, gain=2.0)
    linear("v", embedding_size, 1, gain=1.0)
    return sd


def net_shape(sd):
    channels = sd["conv_in.weight"].shape[0]
    blocks = 0
    while f"resnet.{blocks}.conv_block1.0.weight" in sd:
        blocks += 1
    return channels, blocks, sd["fc1.weight"].shape[0], sd["conv_in.weight"].shape[1]


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.asarray(x))


def _bn(x, sd, prefix):
    return F.batch_norm(x, _t(sd[prefix + ".running_mean"]), _t(sd[prefix + ".running_var"]),
                        _t(sd[prefix + ".weight"]), _t(sd[prefix + ".bias"]), training=False, eps=BN_EPS)


def forward(sd, obs, return_logits=False):
    """obs: float32 [B, C_in, 9, 9] -> (pi softmax [B,81], v tanh [B,1]) as torch tensors."""
    with torch.no_grad():
        x = _t(obs).float()
        _, blocks, _, _ = net_shape(sd)
        x = F.relu(_bn(F.conv2d(x, _t(sd["conv_in.weight"]), _t(sd["conv_in.bias"]), padding=1), sd, "bn_in"))
        for k in range(blocks):                                   # network.py:35-41
            p = f"resnet.{k}."
            y = F.relu(_bn(F.conv2d(x, _t(sd[p + "conv_block1.0.weight"]), _t(sd[p + "conv_block1.0.bias"]),
                                    padding=1), sd, p + "conv_block1.1"))
            y = _bn(F.conv2d(y, _t(sd[p + "conv_block2.0.weight"]), _t(sd[p + "conv_block2.0.bias"]),
                             padding=1), sd, p + "conv_block2.1")
            x = F.relu(y + x)
        x = torch.flatten(x, 1)                                   # network.py:92 (NCHW order)
        x = F.relu(_bn(F.linear(x, _t(sd["fc1.weight"]), _t(sd["fc1.bias"])), sd, "fc_bn1"))
        logits = F.linear(x, _t(sd["pi.weight"]), _t(sd["pi.bias"]))
        v = torch.tanh(F.linear(x, _t(sd["v.weight"]), _t(sd["v.bias"])))
        pi = F.softmax(logits, dim=1)                             # network.py:99 — unmasked
        if return_logits:
            return pi, v, logits
        return pi, v


class OracleNet:
    """`predict` with the reference's signature (nn_wrapper.py:125-136)."""

    def __init__(self, sd):
        self.sd = {k: _t(v) for k, v in sd.items()}
        self.rows = 0

    def predict(self, obs):
        pi, v = forward(self.sd, obs)
        self.rows += int(pi.shape[0])
        return pi.numpy(), v.numpy()


def flops_per_eval(channels, blocks, emb, in_planes=4):
    """F = 2*81*9*C*(C_in + 2*B*C) + 2*81*C*E + 2*E*82."""
    return 2 * 81 * 9 * channels * (in_planes + 2 * blocks * channels) + 2 * 81 * channels * emb + 2 * emb * 82
