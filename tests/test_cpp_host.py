"""The compiled host side: include/eccbatch.hpp (C++17 mirror of the reference's per-curve API over the
C ABI) and its test program tests/cpp/test_host_mirror.cpp, which reads like the reference's own test
modules (RFC 7748 / 8032 / 6979 vectors, NIST KATs, BLS12-381 G1 serialization KATs, error cases).

CPU: the program compiles against the header, links libeccbatch.so, and without a device the
constructor throws (no CPU fallback).  GPU: every check passes on cuda:0."""
import hashlib
import os
import subprocess

import pytest

from oracle import pyref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
EXE = os.path.join(CPP, "_build", "test_host_mirror")


@pytest.fixture(scope="module")
def exe():
    subprocess.check_call(["make", "-s", "-C", CPP])
    return EXE


def write_vectors(golden, path):
    rows = []
    x = golden["x25519"]
    for v in x["rfc7748_5_2"]:
        rows.append(("x25519", v["k"], v["u"], v["r"]))
    rows.append(("x25519_dh", x["dh_6_1"]["a"], x["dh_6_1"]["b"], x["dh_6_1"]["shared"]))
    for v in golden["x448"]["rfc7748_5_2"]:
        rows.append(("x448", v["k"], v["u"], v["r"]))
    for v in golden["ed25519_rfc8032"]:
        rows.append(("ed25519", v["seed"], v["public"], v["message"] or "-", v["signature"]))
    for s in golden["ed25519_edge_scalars_u64"][:12]:
        rows.append(("ed_scalar", int(s).to_bytes(32, "little").hex()))
    for name in ("nist_p256", "nist_p384"):
        for v in golden[name][:24]:
            rows.append((name, v["k"], v["x"], v["y"]))
    e = golden["ecdsa_rfc6979"]["p256r1"]
    for kat in e["kats"]:
        dg = hashlib.new(kat["alg"], kat["message"].encode()).digest()
        z = R.ecdsa_digest_to_scalar(R.P256, dg).hex()
        rows.append(("ecdsa_p256", e["d"], kat["k"], z, kat["r"].rjust(64, "0"), kat["s"].rjust(64, "0")))
    for v in golden["bls12_381_g1"]["compressed"]:
        rows.append(("bls_compressed", v["k"].to_bytes(32, "big").hex(), v["bytes"]))
    for v in golden["bls12_381_g1"]["off_subgroup"]:
        rows.append(("bls_off_subgroup", v["compressed"], v["uncompressed"]))
    with open(path, "w") as f:
        for r in rows:
            f.write(" ".join(r) + "\n")
    return len(rows)


def test_cpp_host_compiles_links_and_refuses_to_run_without_a_device(exe, golden, tmp_path):
    import torch

    assert os.path.exists(exe)
    assert write_vectors(golden, tmp_path / "v.txt") > 60
    r = subprocess.run([exe, "--no-device-check"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    if not torch.cuda.is_available():
        assert "BatchError code -1" in r.stdout


@pytest.mark.gpu
def test_cpp_host_passes_the_reference_vectors_on_the_gpu(exe, golden, tmp_path):
    vec = tmp_path / "v.txt"
    write_vectors(golden, vec)
    r = subprocess.run([exe, str(vec)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout + r.stderr
