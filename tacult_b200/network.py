"""Learner-side torch module with the reference network's parameter names and arithmetic
(`_UtacNNet`, /root/reference/src/tacult/network.py:44-102; `ResidualBlock` 20-41), so that a reference
`state_dict` loads here and a `state_dict` trained here loads into the reference and into the engine
(`tc_load_weights`, BN folded there).

This module exists for the training graph only (autograd); self-play never runs it — leaf evaluation is the
CUDA path behind `Engine.load_weights`.  Parameter names: conv_in, bn_in, resnet.<k>.conv_block1.{0,1},
resnet.<k>.conv_block2.{0,1}, fc1, fc_bn1, pi, v.
"""
import torch
from torch import nn
from torch.nn import functional as F

BOARD_CELLS = 81
IN_PLANES = 4


def _conv_bn(channels_in, channels_out, relu):
    layers = [nn.Conv2d(channels_in, channels_out, kernel_size=3, padding=1), nn.BatchNorm2d(channels_out)]
    if relu:
        layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class _Block(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv_block1 = _conv_bn(channels, channels, relu=True)
        self.conv_block2 = _conv_bn(channels, channels, relu=False)

    def forward(self, x):
        return F.relu(self.conv_block2(self.conv_block1(x)) + x)


class TrunkNet(nn.Module):
    def __init__(self, channels=16, num_residual_blocks=4, embedding_size=256, dropout=0.3, in_planes=IN_PLANES):
        super().__init__()
        self.conv_in = nn.Conv2d(in_planes, channels, kernel_size=3, padding=1)
        self.bn_in = nn.BatchNorm2d(channels)
        self.resnet = nn.Sequential(*[_Block(channels) for _ in range(num_residual_blocks)])
        self.fc1 = nn.Linear(BOARD_CELLS * channels, embedding_size)
        self.fc_bn1 = nn.BatchNorm1d(embedding_size)
        self.pi = nn.Linear(embedding_size, BOARD_CELLS)
        self.v = nn.Linear(embedding_size, 1)
        self.dropout = float(dropout)

    def forward(self, obs):
        """obs f32[B,4,9,9] -> (unmasked soft-max policy f32[B,81], tanh value f32[B,1])."""
        x = F.relu(self.bn_in(self.conv_in(obs)))
        x = self.resnet(x)
        x = F.relu(self.fc_bn1(self.fc1(torch.flatten(x, 1))))        # NCHW flatten: k = c*81 + cell
        x = F.dropout(x, p=self.dropout, training=self.training)
        return F.softmax(self.pi(x), dim=1), torch.tanh(self.v(x))

    @classmethod
    def from_state_dict(cls, state_dict, dropout=0.3):
        """Architecture read off the tensor shapes (as weights.state_dict_to_blob does)."""
        from .weights import _strip
        state_dict = _strip(state_dict)          # torch.compile'd reference checkpoints prefix "_orig_mod."
        channels = int(state_dict["conv_in.weight"].shape[0])
        in_planes = int(state_dict["conv_in.weight"].shape[1])
        blocks = 1 + max(int(k.split(".")[1]) for k in state_dict if k.startswith("resnet."))
        emb = int(state_dict["fc1.weight"].shape[0])
        net = cls(channels, blocks, emb, dropout, in_planes)
        net.load_state_dict({k: torch.as_tensor(v) for k, v in state_dict.items()})
        return net
