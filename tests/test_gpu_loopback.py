"""The multi-rank schedule on ONE GPU (tests/loopback_worker.py): G contexts of one process as the ranks of a box.  This is
the N > 1 parity evidence a one-GPU box can produce — tests/test_gpu_multi.py repeats it across real GPUs over NVLink when the box
has them.  Each case runs in its own process with one hardware queue per stream (CUDA_DEVICE_MAX_CONNECTIONS=32) and eager module
loading (CUDA_MODULE_LOADING=EAGER)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(which, timeout=600, **env):
    # one hardware queue per stream, and every kernel loaded up front: with lazy module loading the FIRST launch of a kernel waits
    # for the device to drain — forever, if what runs there is a barrier waiting for a rank this very host thread has yet to enqueue
    e = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32", CUDA_MODULE_LOADING="EAGER", **env)
    r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "loopback_worker.py"), which], capture_output=True, text=True, timeout=timeout, env=e)
    assert r.returncode == 0 and f"OK [{which}]" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_pipelined_frames_equal_one_stream_frames():
    """one context: eight frames back to back through the three-stream pipeline (double-buffered texture sets) == F184_FLAG_NO_OVERLAP"""
    run("single")


def test_two_loopback_ranks_equal_one_context():
    """volumes, both texture sets, image rows; the level-0 skip of the gather incl. its rough -> glossy transitions; pipelined frames"""
    run("box2")


def test_four_loopback_ranks_equal_one_context():
    run("box4")


def test_gather_with_many_units_per_cta():
    """three CTAs instead of 296 make every gather CTA walk many 64-brick units (what a 1024^3 volume does to the full grid)"""
    run("sponza", F184_GATHER_CTAS="3")


def test_full_fragment_queues_fall_back_to_remote_reductions():
    """a receive queue of 2000 records overflows at once: everything beyond it takes the system-scope reduction path, same volume"""
    run("box2", F184_FRAG_QUEUE_RECORDS="2000")


def test_four_loopback_ranks_sponza_256():
    run("sponza")


def test_barrier_with_a_missing_peer_is_an_error():
    run("timeout", F184_BARRIER_TIMEOUT_MS="300")


def test_static_cache_equals_full_voxelization():
    """static / dynamic split (f184_static_cache_capture): one context (synchronised and pipelined frames, re-capture, clear, camera
    check) and two loopback ranks, bit for bit against voxelizing everything every frame"""
    run("static")
