"""SDXL UNet graph -- same module tree and parameter names as the reference's
diffusers_rewrite/sdxl.py (:505-556), executed by dgq_b200.engine."""
import torch.nn as nn

from .common import (Attention, Downsample2D, FeedForward, GEGLU, ResnetBlock2D, TimestepEmbedding,
                     Timesteps, Upsample2D, BasicTransformerBlockBase, Transformer2DModelBase, _BlockList,
                     UNetBase)
from .. import engine

__all__ = ["Timesteps", "TimestepEmbedding", "ResnetBlock2D", "Attention", "GEGLU", "FeedForward",
           "BasicTransformerBlock", "Transformer2DModel", "Downsample2D", "Upsample2D", "DownBlock2D",
           "CrossAttnDownBlock2D", "CrossAttnUpBlock2D", "UpBlock2D", "UNetMidBlock2DCrossAttn",
           "UNet2DConditionModel"]


class BasicTransformerBlock(BasicTransformerBlockBase):
    def __init__(self, hidden_size):
        super().__init__(hidden_size, 2048, num_heads=None)  # head_dim 64


class Transformer2DModel(Transformer2DModelBase):
    def __init__(self, in_channels, out_channels, n_layers):
        super().__init__()
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-06, affine=True)
        self.proj_in = nn.Linear(in_channels, out_channels, bias=True)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(out_channels) for _ in range(n_layers)])
        self.proj_out = nn.Linear(out_channels, out_channels, bias=True)


class DownBlock2D(_BlockList):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels, out_channels, conv_shortcut=False),
                                      ResnetBlock2D(out_channels, out_channels, conv_shortcut=False)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels, out_channels)])


class CrossAttnDownBlock2D(_BlockList):
    def __init__(self, in_channels, out_channels, n_layers, has_downsamplers=True):
        super().__init__()
        self.attentions = nn.ModuleList([Transformer2DModel(out_channels, out_channels, n_layers),
                                         Transformer2DModel(out_channels, out_channels, n_layers)])
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels, out_channels),
                                      ResnetBlock2D(out_channels, out_channels, conv_shortcut=False)])
        self.downsamplers = None
        if has_downsamplers:
            self.downsamplers = nn.ModuleList([Downsample2D(out_channels, out_channels)])


class CrossAttnUpBlock2D(_BlockList):
    def __init__(self, in_channels, out_channels, prev_output_channel, n_layers):
        super().__init__()
        self.attentions = nn.ModuleList([Transformer2DModel(out_channels, out_channels, n_layers) for _ in range(3)])
        self.resnets = nn.ModuleList([ResnetBlock2D(prev_output_channel + out_channels, out_channels),
                                      ResnetBlock2D(2 * out_channels, out_channels),
                                      ResnetBlock2D(out_channels + in_channels, out_channels)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, out_channels)])


class UpBlock2D(_BlockList):
    def __init__(self, in_channels, out_channels, prev_output_channel):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(out_channels + prev_output_channel, out_channels),
                                      ResnetBlock2D(out_channels * 2, out_channels),
                                      ResnetBlock2D(out_channels + in_channels, out_channels)])


class UNetMidBlock2DCrossAttn(_BlockList):
    def __init__(self, in_features):
        super().__init__()
        self.attentions = nn.ModuleList([Transformer2DModel(in_features, in_features, n_layers=10)])
        self.resnets = nn.ModuleList([ResnetBlock2D(in_features, in_features, conv_shortcut=False),
                                      ResnetBlock2D(in_features, in_features, conv_shortcut=False)])


class UNet2DConditionModel(UNetBase):
    def __init__(self):
        super().__init__()
        self.register_to_config(in_channels=4, addition_time_embed_dim=256, sample_size=128,
                                time_cond_proj_dim=None)
        self.conv_in = nn.Conv2d(4, 320, kernel_size=3, stride=1, padding=1)
        self.time_proj = Timesteps()
        self.time_embedding = TimestepEmbedding(in_features=320, out_features=1280)
        self.add_time_proj = Timesteps(256)
        self.add_embedding = TimestepEmbedding(in_features=2816, out_features=1280)
        self.down_blocks = nn.ModuleList([
            DownBlock2D(in_channels=320, out_channels=320),
            CrossAttnDownBlock2D(in_channels=320, out_channels=640, n_layers=2),
            CrossAttnDownBlock2D(in_channels=640, out_channels=1280, n_layers=10, has_downsamplers=False)])
        self.up_blocks = nn.ModuleList([
            CrossAttnUpBlock2D(in_channels=640, out_channels=1280, prev_output_channel=1280, n_layers=10),
            CrossAttnUpBlock2D(in_channels=320, out_channels=640, prev_output_channel=1280, n_layers=2),
            UpBlock2D(in_channels=320, out_channels=320, prev_output_channel=640)])
        self.mid_block = UNetMidBlock2DCrossAttn(1280)
        self.conv_norm_out = nn.GroupNorm(32, 320, eps=1e-05, affine=True)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(320, 4, kernel_size=3, stride=1, padding=1)

    def forward(self, sample, timesteps, encoder_hidden_states, added_cond_kwargs, **kwargs):
        return [engine.unet_forward(self, sample, timesteps, encoder_hidden_states, added_cond_kwargs)]
