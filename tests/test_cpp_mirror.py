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
                 "get_global_rank", "get_local_idx", "group_by_batch", "distribute_table", "collect_tables",
                 "nccl_comm"]:
        assert name in text, name


def test_communication_group_semantics():
    """CPU: CommunicationGroup(grid, stride) membership, the reference's documented example
    (src/all_to_all_comm.hpp:72-113): 16 ranks, grid 8, stride 2 -> {0,2,4,6} {1,3,5,7} {8,10,12,14} {9,11,13,15}."""
    exe = os.path.join(BIN, "test_comm_group")
    if not os.path.exists(exe):
        pytest.skip("bin/test_comm_group not built")

    def group(rank, grid, stride):
        out = subprocess.run([exe, str(grid), str(stride)], env=dict(os.environ, RANK=str(rank), WORLD_SIZE="16"),
                             capture_output=True, text=True, check=True).stdout.split()
        return int(out[0]), int(out[1]), [int(x) for x in out[2:]]

    want = {0: [0, 2, 4, 6], 1: [1, 3, 5, 7], 8: [8, 10, 12, 14], 9: [9, 11, 13, 15]}
    for rank in range(16):
        size, idx, members = group(rank, 8, 2)
        head = rank // 8 * 8 + rank % 2
        assert size == 4 and members == want[head] and members[idx] == rank
    # stride 1 over all ranks (what shuffle_on / the NVLink stage use)
    size, idx, members = group(5, 16, 1)
    assert size == 16 and idx == 5 and members == list(range(16))
    # the IB-stage group of distributed_inner_join: CommunicationGroup(N, G) with N=8, G=2
    size, idx, members = group(5, 8, 2)
    assert members == [1, 3, 5, 7] and idx == 2


def test_bootstrap_rendezvous_two_ranks(tmp_path):
    """CPU: the MPI-free ncclUniqueId broadcast (host/bootstrap.cpp): ranks started by one parent agree on
    rank 0's bytes; a stale file of an earlier launch on the same port is never picked up (per-launch nonce)."""
    exe = os.path.join(BIN, "test_bootstrap")
    if not os.path.exists(exe):
        pytest.skip("bin/test_bootstrap not built")
    for salt in ("A", "B"):  # two launches on the same port; each has its own parent, like two torchrun agents
        env = dict(os.environ, WORLD_SIZE="2", MASTER_PORT="29999", DJ_RENDEZVOUS_DIR=str(tmp_path))
        r = subprocess.run(["sh", "-c", f"RANK=1 {exe} {salt} & RANK=0 {exe} {salt}; wait"], env=env,
                           capture_output=True, text=True, timeout=60)
        outs = sorted(line.split() for line in r.stdout.strip().splitlines())
        assert r.returncode == 0 and len(outs) == 2, r.stdout + r.stderr
        assert outs[0][0] == "0" and outs[1][0] == "1" and outs[0][2] == outs[1][2]


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
def test_compare_against_single_gpu_single_rank():
    """Port of test/compare_against_single_gpu.cu (G3) with distribute_table / collect_tables; on one rank
    the distributed path must equal the plain single-GPU join for the whole type / odf matrix."""
    r = _run(1, "compare_against_single_gpu")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'Test case "compare_against_single_gpu" passes successfully.' in r.stderr
    assert r.stderr.count("passes successfully") >= 20 and "FAILED" not in r.stderr


@pytest.mark.gpu
def test_shuffle_on_benchmark_verifies_its_result():
    """benchmark/shuffle_on (config 4 at a small size) checks rows, co-location and the multiset checksum."""
    r = _run(1, "shuffle_on", "--nrows", "3000000", "--iterations", "1")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "shuffle_on check: OK" in r.stdout


@pytest.mark.gpu
def test_cpp_programs_multi_rank():
    import torch

    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    for exe in ("compare_against_analytical", "test_shuffle_on", "compare_against_single_gpu"):
        r = _run(n, exe)
        assert r.returncode == 0, exe + r.stdout[-2000:] + r.stderr[-2000:]
        assert "passes successfully" in r.stderr
