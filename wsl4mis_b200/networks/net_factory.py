"""net_factory with the reference's signature and behaviour (networks/net_factory.py:6-22): returns a CUDA
module for the accelerated model names, ``None`` for unknown names."""
from .unet import UNet, UNet_CCT


def net_factory(net_type="unet", in_chns=1, class_num=3):
    if net_type == "unet":
        return UNet(in_chns=in_chns, class_num=class_num).cuda()
    if net_type == "unet_cct":
        return UNet_CCT(in_chns=in_chns, class_num=class_num).cuda()
    if net_type in ("unet_cct_3h", "unet_ds", "efficient_unet", "pnet"):
        raise NotImplementedError(f"net_type '{net_type}' is outside the accelerated hot path (unet, unet_cct)")
    return None
