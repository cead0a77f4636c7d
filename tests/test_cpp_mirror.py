"""The C++ mirror of the reference API (distributed-join_b200/host) exercised through the ported
reference programs: test/compare_against_analytical (G1) and test/test_shuffle_on (G4)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "distributed-join_b200", "bin")


def test_cpp_headers_mirror_reference_names():
    """CPU: the mirror declares the reference's public entry points with the same names."""
    host = os.path.join(ROOT, "distributed-join_b200", "host")
    text = "".join(open(os.path.join(host, f)).read() for f in os.listdir(host) if f.endswith((".hpp",)))
    for name in ["distributed_inner_join", "shuffle_on", "all_to_all_comm", "postprocess_all_to_all_comm",
                 "communicate_sizes", "append_to_all_to_all_comm_buffers", "warmup_all_to_all",
                 "class AllToAllCommunicator", "class CommunicationGroup", "struct AllToAllCommBuffer",
                 "class Communicator", "class NCCLCommunicator", "set_cuda_device",
                 "setup_memory_pool_and_communicator", "destroy_memory_pool_and_communicator",
                 "generate_tables_distributed", "generate_build_probe_tables", "ColumnCompressionOptions",
                 "generate_none_compression_options", "allocate_communicated_table", "launch_communication",
                 "get_global_rank", "get_local_idx", "group_by_batch"]:
        assert name in text, name


def _run(nproc, exe, *args):
    if nproc == 1:
        return subprocess.run([os.path.join(BIN, exe), *args], cwd=ROOT, capture_output=True, text=True, timeout=300,
                              env=dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0"))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1",
                           f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1", "--master-port", str(port),
                           os.path.join(BIN, exe), *args], cwd=ROOT, capture_output=True, text=True, timeout=600)


@pytest.mark.gpu
def test_compare_against_analytical_single_rank():
    r = _run(1, "compare_against_analytical")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'Test case "compare_against_analytical" passes successfully.' in r.stderr


@pytest.mark.gpu
def test_cpp_programs_multi_rank():
    import torch

    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    for exe in ("compare_against_analytical", "test_shuffle_on"):
        r = _run(n, exe)
        assert r.returncode == 0, exe + r.stdout[-2000:] + r.stderr[-2000:]
        assert "passes successfully" in r.stderr
