from .dcn import CONV_LAYERS, ModulatedDeformConv2dPackPlugin, ModulatedDeformConv2dPackPlugin2

__all__ = ["CONV_LAYERS", "ModulatedDeformConv2dPackPlugin", "ModulatedDeformConv2dPackPlugin2"]
