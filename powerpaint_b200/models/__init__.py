from .unet_2d_condition import (BrushNetModel, BrushNetOutput, ControlNetModel, ControlNetOutput,
                                UNet2DConditionModel, UNet2DConditionOutput)
from .spec import param_shapes, synthetic_state_dict

__all__ = ["BrushNetModel", "BrushNetOutput", "ControlNetModel", "ControlNetOutput", "UNet2DConditionModel",
           "UNet2DConditionOutput", "param_shapes", "synthetic_state_dict"]
