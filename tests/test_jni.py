"""The JNI half of the drop-in boundary without a JDK: vpca_jni.c must type-check against the stub jni.h with
-Wall -Wextra -Werror, and its argument validation must reject bad lengths / nulls before touching the library
(run against the mock JNIEnv of tests/jni_harness.c; no GPU needed).  The GPU run of the shim is in
tests/test_pool_gpu.py."""
import json
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_jni_shim_type_checks_against_stub_header():
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    proc = subprocess.run([gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", f"-I{ROOT / 'tests' / 'stubs'}",
                           f"-I{ROOT / 'include'}", str(ROOT / "spark_examples_b200" / "jvm" / "vpca_jni.c")],
                          capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr


def test_jni_shim_uses_no_critical_sections():
    """JNI critical regions must not block, and every vpca_* call may: the shim copies regions or takes
    direct buffers instead."""
    text = (ROOT / "spark_examples_b200" / "jvm" / "vpca_jni.c").read_text()
    code = "\n".join(line for line in text.splitlines() if not line.lstrip().startswith(("*", "/*")))
    assert "PrimitiveArrayCritical" not in code
    assert "GetArrayLength" in code and "GetDirectBufferCapacity" in code


def test_jni_shim_validates_arguments_before_calling_the_library():
    import __graft_entry__ as entry
    exe = ROOT / "tests" / "_build" / "jni_harness"
    if not exe.exists():
        entry.build()
    proc = subprocess.run([str(exe), "validate"], capture_output=True, text=True, timeout=120)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert json.loads(proc.stdout.strip().splitlines()[-1])["failures"] == 0


def test_scala_binding_declares_what_the_shim_exports():
    """Every `@native def` of NativePca / NativePcaPool has its JNI symbol in vpca_jni.c and vice versa."""
    import re
    shim = (ROOT / "spark_examples_b200" / "jvm" / "vpca_jni.c").read_text()
    for cls, macro in (("NativePca", "PCA"), ("NativePcaPool", "POOL")):
        scala = (ROOT / "spark_examples_b200" / "jvm" / f"{cls}.scala").read_text()
        declared = set(re.findall(r"@native def (\w+)\(", scala))
        exported = set(re.findall(r"JNICALL %s\((\w+)\)" % macro, shim))
        assert declared == exported, (cls, declared ^ exported)
