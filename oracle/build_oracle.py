"""TEST INFRASTRUCTURE ONLY. Compiles oracle/structural_oracle.c -> oracle/_build/liboracle.so (gcc)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "liboracle.so")


def build_oracle(force=False):
    src = os.path.join(HERE, "structural_oracle.c")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", OUT, src, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build_oracle(force=True))
