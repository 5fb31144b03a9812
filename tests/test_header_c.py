"""include/clsr_b200.h is a C header: it must compile as C11 with nothing but <stdint.h>."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "clsr_b200.h"\nint main(void) { clsr_config c; clsr_batch b; (void)c; (void)b; '
                   'return sizeof(clsr_losses) == 36 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0
