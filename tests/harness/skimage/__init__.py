"""stand-in: dataloaders/utils.py imports skimage.measure at module level; no WSS script calls into it"""
