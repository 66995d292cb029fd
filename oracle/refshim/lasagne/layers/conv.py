from . import (BaseConvLayer, Conv2DLayer, TransposedConv2DLayer, Deconv2DLayer, DilatedConv2DLayer,   # noqa: F401
               conv_output_length, conv_input_length)
