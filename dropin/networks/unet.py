from wsl4mis_b200.networks.unet import *  # noqa: F401,F403
from wsl4mis_b200.networks.unet import (ConvBlock, DownBlock, UpBlock, Encoder, Decoder, Decoder_DS, Decoder_URDS, Dropout,  # noqa: F401
                                        FeatureDropout, FeatureNoise, UNet, UNet_DS, UNet_CCT, UNet_CCT_3H)
