"""The import shim must win from the directory the unchanged scripts run in (`sys.path[0]` is the script
directory `code/`, ahead of PYTHONPATH; namespace-package merging used to hand the scripts the reference's own modules)."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROBE = textwrap.dedent("""
    import json, sys
    import utils.losses, utils.gate_crf_loss, utils.ramps, utils.metrics
    import networks.net_factory, networks.unet, networks.pnet
    import dataloaders.dataset, val_2D
    from utils import losses, metrics, ramps
    from dataloaders.dataset import BaseDataSets, RandomGenerator
    from val_2D import test_single_volume, test_single_volume_cct, test_single_volume_ds
    print(json.dumps({m: sys.modules[m].__file__ for m in ("utils.losses", "utils.gate_crf_loss", "utils.ramps", "utils.metrics",
          "networks.net_factory", "networks.unet", "networks.pnet", "dataloaders.dataset", "val_2D")}))
""")


def _code_dir(tmp_path):
    """a code/-like directory: the staged reference files when present, otherwise stand-ins with the same names"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import build_ref
        code = build_ref.extract(str(tmp_path))
    finally:
        sys.path.pop(0)
    if code is None:
        code = tmp_path / "code"
        for rel in ("utils/losses.py", "utils/gate_crf_loss.py", "utils/ramps.py", "utils/metrics.py", "networks/unet.py",
                    "networks/net_factory.py", "networks/pnet.py", "dataloaders/dataset.py", "val_2D.py"):
            f = code / rel
            f.parent.mkdir(parents=True, exist_ok=True)
            f.write_text("ORIGIN = 'reference stand-in'\n")
        code = str(code)
    return code


def test_shim_wins_from_the_script_directory(tmp_path):
    code = _code_dir(tmp_path)
    probe = os.path.join(code, "probe_imports.py")
    open(probe, "w").write(PROBE)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT, os.path.join(ROOT, "tests", "harness")])
    r = subprocess.run([sys.executable, probe], cwd=code, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    where = json.loads(r.stdout.strip().splitlines()[-1])
    for m in ("utils.losses", "utils.gate_crf_loss", "utils.ramps", "networks.net_factory", "networks.unet", "networks.pnet",
              "dataloaders.dataset", "val_2D"):
        assert where[m].startswith(os.path.join(ROOT, "dropin")), (m, where[m])
    for m in ("utils.metrics",):                          # not replaced: still the file next to the script
        assert where[m].startswith(code), (m, where[m])


def test_launcher_sets_the_same_resolution(tmp_path):
    code = _code_dir(tmp_path)
    probe = os.path.join(code, "probe_imports.py")
    open(probe, "w").write(PROBE)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([ROOT, os.path.join(ROOT, "tests", "harness")])
    r = subprocess.run([sys.executable, "-m", "wsl4mis_b200.run", probe], cwd=code, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    where = json.loads(r.stdout.strip().splitlines()[-1])
    assert where["utils.losses"].startswith(os.path.join(ROOT, "dropin")) and where["utils.metrics"].startswith(code)


def test_script_signature_dataset_on_the_host():
    """BaseDataSets(base_dir, split, transform, fold, sup_type) as train_weakly_supervised_pCE_GatedCRFLoss_2D.py:75-80 calls it,
    iterated by a DataLoader with worker processes (host-only samples)."""
    import torch
    from torchvision import transforms
    from wsl4mis_b200.dataloaders.dataset import BaseDataSets, RandomGenerator
    db = BaseDataSets(base_dir="synthetic:24", split="train", transform=transforms.Compose([RandomGenerator([256, 256])]),
                      fold="fold1", sup_type="scribble")
    val = BaseDataSets(base_dir="synthetic:24", fold="fold1", split="val")
    assert len(db) > 16 and len(val) == 20
    loader = torch.utils.data.DataLoader(db, batch_size=4, shuffle=True, num_workers=2)
    b = next(iter(loader))
    assert b["image"].shape == (4, 1, 256, 256) and b["image"].dtype == torch.float32
    assert b["label"].shape == (4, 256, 256) and b["label"].dtype == torch.uint8
    assert set(b["label"].unique().tolist()) <= {0, 1, 2, 3, 4} and (b["label"] == 4).float().mean() > 0.9
    v = val[0]
    assert v["image"].ndim == 3 and v["label"].shape == v["image"].shape and v["label"].max() <= 3
    assert b["idx"][0].startswith("patient")
