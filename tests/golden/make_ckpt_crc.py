"""Extract the per-tensor masked crc32c values TensorFlow stored in the reference's shipped checkpoint
(examples/00_quick_start/CLSR/taobao-clsr-debug/model.tar.gz, epoch_3.index) into ckpt_crc.json.
They are TF-written known answers for clsr_b200.tf_bundle.crc32c / _mask (tests/test_abi.py): the dense
variables are stored whole in ckpt_slice.npz, so the test recomputes their checksums from the bytes.
Run (build container only, needs /root/reference):  python tests/golden/make_ckpt_crc.py
"""
import json
import os
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from clsr_b200 import tf_bundle as tb  # noqa: E402

tmp = tempfile.mkdtemp()
with tarfile.open("/root/reference/examples/00_quick_start/CLSR/taobao-clsr-debug/model.tar.gz") as t:
    t.extractall(tmp)
prefix = os.path.join(tmp, "model", "epoch_3")
ent = tb.read_index(prefix)
full = tb.read_bundle(prefix, verify_crc=True)   # every tensor of the TF-written bundle passes our check
out = {k: {"crc32c_masked": int(v[5]), "bytes": int(v[4])} for k, v in ent.items()}
with open(os.path.join(HERE, "ckpt_crc.json"), "w") as f:
    json.dump(out, f, indent=0, sort_keys=True)
print("verified and wrote %d entries" % len(out), sum(v.nbytes for v in full.values()), "bytes")
