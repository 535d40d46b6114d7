"""Model factory with the reference's call signature ``net_factory(net_type, in_chns, class_num)`` and its
behaviour at the edges (networks/net_factory.py:6-22): the model is returned already on the GPU, an unknown name yields
``None``.  The architectures the executors plan (unet, unet_cct, unet_ds, unet_cct_3h, pnet) are constructible here."""
from .pnet import PNet2D
from .unet import UNet, UNet_CCT, UNet_CCT_3H, UNet_DS

_ACCELERATED = {"unet": UNet, "unet_cct": UNet_CCT, "unet_ds": UNet_DS, "unet_cct_3h": UNet_CCT_3H}
_KNOWN_BUT_OFF_PATH = ("efficient_unet",)


def net_factory(net_type="unet", in_chns=1, class_num=3):
    ctor = _ACCELERATED.get(net_type)
    if ctor is not None:
        return ctor(in_chns=in_chns, class_num=class_num).cuda()
    if net_type == "pnet":                      # reference net_factory.py:18-19
        return PNet2D(in_chns, class_num, 64, [1, 2, 4, 8, 16]).cuda()
    if net_type in _KNOWN_BUT_OFF_PATH:
        raise NotImplementedError(f"net_type '{net_type}' is outside the accelerated hot path ({', '.join(_ACCELERATED)})")
    return None
