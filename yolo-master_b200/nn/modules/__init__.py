"""Operator library mirroring `ultralytics.nn.modules` (same class names / signatures / state_dict keys)."""
from .block import A2C2f, AAttn, ABlock, Attention, Bottleneck, C2f, C2PSA, C3, C3k, C3k2, PSABlock, SPPF
from .conv import Concat, Conv, DWConv, PlainConv2d, Upsample, autopad
from .gated import (AdaptiveGateMoE, ContextRefinedLowRankHybridAdaptiveGateMoE, DetailAwareLowRankHybridAdaptiveGateMoE,
                    DualStreamGateRouter, FusedAdaptiveGateMoE, FusedExpertGroup, HybridAdaptiveGateMoE, LowRankFusedExpertGroup,
                    LowRankHybridAdaptiveGateMoE, PyramidContextMixer, RefinedLowRankHybridAdaptiveGateMoE, SharedInvertedExpertGroup,
                    UltimateOptimizedMoE, VisualDetailGate, VisualEnhancedAdaptiveGateMoE, ZeroCostRouter, DualStreamGateRouterV2,
                    HybridAdaptiveGateMoEv2, OptimalHybridGateMoE, MultiHeadRouterMoE, GatedFusionMoE, MultiHeadRouterV3, CrossPathGate,
                    SharedExpertMoE, DiversifiedExpertGroup, DiversifiedExpertMoE)
from .head import DFL, OBB, Classify, Detect, Pose, Proto, Segment
from .latent import DenseChannelExpert, LatentMixture, LatentRouter
from .moa import C2fMoA, MoABlock
from .mot import C2fMoT, MoTBlock
from .moe import (A2C2fMoE, ABlockMoE, DepthwiseSeparableConv, DynamicRoutingLayer, EfficientExpertGroup, EfficientSpatialRouter,
                  ES_MOE, OptimizedMOEImproved, SimpleExpert, UltraEfficientRouter, UltraOptimizedMoE, get_safe_groups)

ModularRouterExpertMoE = OptimizedMOEImproved   # alias of the reference (moe/modules.py:1745), the block of the v0_1 zoo

__all__ = (
    "ModularRouterExpertMoE", "UltraOptimizedMoE", "UltraEfficientRouter", "LatentMixture", "LatentRouter", "DenseChannelExpert",
    "Conv", "DWConv", "Concat", "Upsample", "PlainConv2d", "autopad",
    "Bottleneck", "C2f", "C3", "C3k", "C3k2", "SPPF", "Attention", "PSABlock", "C2PSA", "AAttn", "ABlock", "A2C2f",
    "EfficientSpatialRouter", "SimpleExpert", "OptimizedMOEImproved", "ABlockMoE", "A2C2fMoE", "get_safe_groups",
    "DynamicRoutingLayer", "DepthwiseSeparableConv", "EfficientExpertGroup", "ES_MOE",
    "Detect", "DFL", "Pose", "Proto", "Segment", "OBB", "Classify", "C2fMoT", "MoTBlock", "C2fMoA", "MoABlock",
    "AdaptiveGateMoE", "FusedAdaptiveGateMoE", "HybridAdaptiveGateMoE", "LowRankHybridAdaptiveGateMoE",
    "RefinedLowRankHybridAdaptiveGateMoE", "DetailAwareLowRankHybridAdaptiveGateMoE", "ContextRefinedLowRankHybridAdaptiveGateMoE",
    "VisualEnhancedAdaptiveGateMoE", "UltimateOptimizedMoE", "ZeroCostRouter", "DualStreamGateRouterV2", "HybridAdaptiveGateMoEv2",
    "OptimalHybridGateMoE", "MultiHeadRouterMoE", "GatedFusionMoE", "SharedExpertMoE", "DiversifiedExpertMoE", "DiversifiedExpertGroup", "MultiHeadRouterV3", "CrossPathGate", "DualStreamGateRouter", "FusedExpertGroup", "LowRankFusedExpertGroup", "SharedInvertedExpertGroup",
    "VisualDetailGate", "PyramidContextMixer",
)
