"""stand-in package: only medpy.metric.binary.{dc, hd95} are used by the reference (utils/metrics.py:2, val_2D.py:3)"""
