"""Model factory with the reference's call signature ``net_factory(net_type, in_chns, class_num)`` and its
behaviour at the edges (networks/net_factory.py:6-22): the model is returned already on the GPU, an unknown name yields
``None``.  Only the architectures the executor plans (unet, unet_cct, unet_ds) are constructible here."""
from .unet import UNet, UNet_CCT, UNet_CCT_3H, UNet_DS

_ACCELERATED = {"unet": UNet, "unet_cct": UNet_CCT, "unet_ds": UNet_DS, "unet_cct_3h": UNet_CCT_3H}
_KNOWN_BUT_OFF_PATH = ("efficient_unet",)


def net_factory(net_type="unet", in_chns=1, class_num=3):
    ctor = _ACCELERATED.get(net_type)
    if ctor is not None:
        return ctor(in_chns=in_chns, class_num=class_num).cuda()
    if net_type in _KNOWN_BUT_OFF_PATH:
        raise NotImplementedError(f"net_type '{net_type}' is outside the accelerated hot path ({', '.join(_ACCELERATED)})")
    return None
