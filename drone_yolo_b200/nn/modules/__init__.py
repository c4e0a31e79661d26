"""Modules of the Drone-YOLO graphs (the subset of ultralytics.nn.modules the three model YAMLs name)."""
from .block import DFL, SPPF, Bottleneck, C2f, RepVGGBlock, conv_bn
from .conv import Concat, Conv, DWConv, RepConv, autopad, fold_bn
from .head import Detect

__all__ = ("Conv", "DWConv", "RepConv", "Concat", "DFL", "SPPF", "C2f", "Bottleneck", "RepVGGBlock", "conv_bn",
           "Detect", "autopad", "fold_bn")
