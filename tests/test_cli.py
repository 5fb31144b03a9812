"""The quick-start CLI: (CPU) the reference's own, unmodified sequential.py imports and drives this
package up to model construction; (GPU) this repository's CLI trains, checkpoints, reloads and
evaluates end to end on a synthetic dataset directory."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CLI = "/root/reference/examples/00_quick_start/sequential.py"


@pytest.mark.skipif(not os.path.exists(REF_CLI), reason="reference tree only exists in the build container")
def test_reference_cli_drives_this_package(tmp_path):
    """Run the reference's sequential.py unchanged with this repo's reco_utils + tensorflow shim;
    CLSRModel is intercepted (no GPU here) and must receive the hparams the flags describe."""
    from clsr_b200.synth_dataset import write_dataset
    write_dataset(str(tmp_path / "data" / "taobao"), train_lines=40, valid_users=4, test_users=4)
    qs = tmp_path / "repo" / "examples" / "00_quick_start"
    qs.mkdir(parents=True)
    os.symlink(os.path.join(ROOT, "reco_utils"), tmp_path / "repo" / "reco_utils")
    os.symlink(os.path.join(ROOT, "clsr_b200"), tmp_path / "repo" / "clsr_b200")
    driver = tmp_path / "drive.py"
    driver.write_text(
        "import sys, runpy, json\n"
        "sys.path.insert(0, %r)\n"
        "import reco_utils.recommender.deeprec.models.sequential.clsr as C\n"
        "class Fake:\n"
        "    def __init__(self, hparams, it, seed=None):\n"
        "        self.hp = hparams; self.it = it(hparams, None)\n"
        "    def fit(self, *a, **k):\n"
        "        print('HP', json.dumps({k: v for k, v in self.hp.values().items() if isinstance(v, (int, float, str, bool))}))\n"
        "        print('ITER', type(self.it).__name__, a, sorted(k.items())); raise SystemExit(0)\n"
        "C.CLSRModel = Fake\n"
        "sys.argv = ['sequential.py', '--dataset', 'taobao', '--data_path', %r, '--batch_size', '8', '--epochs', '2']\n"
        "runpy.run_path(%r, run_name='__main__')\n" % (str(tmp_path / "repo"), str(tmp_path / "data"), REF_CLI))
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "clsr_b200", "shims"))
    r = subprocess.run([sys.executable, str(driver)], cwd=str(qs), env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Tensorflow version: 1.15.2-clsr_b200-shim" in r.stdout
    assert "ITER SASequentialIterator" in r.stdout and "'wauc'" in r.stdout
    hp = __import__("json").loads([l for l in r.stdout.splitlines() if l.startswith("HP ")][0][3:])
    assert hp["batch_size"] == 8 and hp["epochs"] == 2 and hp["max_seq_length"] == 50
    assert hp["contrastive_loss"] == "triplet" and hp["embed_l2"] == 1e-6 and hp["is_clip_norm"] == 1


@pytest.mark.gpu
def test_cli_trains_checkpoints_and_evaluates(tmp_path, cuda_lib):
    from clsr_b200.synth_dataset import write_dataset
    from clsr_b200 import tf_bundle as tb
    write_dataset(str(tmp_path / "data" / "taobao"), train_lines=400, valid_users=30, test_users=20, test_num_ngs=9)
    cli = os.path.join(ROOT, "examples", "00_quick_start", "sequential.py")
    base = [sys.executable, cli, "--dataset", "taobao", "--data_path", str(tmp_path / "data"), "--batch_size", "50",
            "--save_path", str(tmp_path / "save"), "--test_num_ngs", "9", "--show_step", "4"]
    r = subprocess.run(base + ["--epochs", "2"], cwd=os.path.dirname(cli), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "eval valid at epoch 1" in r.stdout and "best epoch" in r.stdout and "wauc" in r.stdout
    mdir = str(tmp_path / "save" / "CLSR" / "taobao-clsr-debug" / "model")
    ck = tb.latest_checkpoint(mdir)
    assert ck and os.path.basename(ck).startswith("epoch_")
    names = tb.read_index(ck)
    assert len(names) == 85 and "sequential/clsr/short_term/time4lstm/time4lstm_cell/kernel" in names
    r2 = subprocess.run(base + ["--only_test"], cwd=os.path.dirname(cli), capture_output=True, text=True, timeout=900)
    assert r2.returncode == 0, r2.stdout[-3000:] + r2.stderr[-3000:]
    # the --only_test metrics equal the ones printed after training for the same checkpoint
    last = [l for l in r.stdout.splitlines() if l.startswith("{") and "wauc" in l][-1]
    again = [l for l in r2.stdout.splitlines() if l.startswith("{") and "wauc" in l][-1]
    assert last == again
