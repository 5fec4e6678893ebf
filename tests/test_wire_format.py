"""B2 wire format pinned by executing the reference's own sender (CPU, no GPU): the unmodified
multiple_connections_network_client_sender.c, compiled from where it lies (oracle/Makefile: ref -> oracle/_ref/ref_sender),
runs against a capture server on loopback -- its literal server address is redirected by an LD_PRELOAD shim
(oracle/shim/connect_local.c).  What fr_ingest_* assumes about a sender (include/fleetrec.h, csrc/fr_ingest.cu) is what
the binary does; the Python sender the GPU ingest tests use (tests/wire.py) produces the same stream."""
import os
import subprocess
import threading

import numpy as np
import pytest

import wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SENDER = os.path.join(ROOT, "oracle", "_ref", "ref_sender")
SHIM = os.path.join(ROOT, "oracle", "_ref", "libconnect_local.so")

# GPU/final_network_cublasLt_1_node_no_FIFO_scatter/constant.h:21-42 (baked into the binary)
INPUT_FEATURE_LEN, BATCH_SIZE, THREAD_NUM, PORT = 880, 128, 4, 8080
TOTAL_BATCH_NUM = 2 * 1024 * 1024 // BATCH_SIZE
BLOCK_SIZE = BATCH_SIZE * INPUT_FEATURE_LEN * 4


def _all_ones(block_u8):
    return bool(np.all(block_u8.view("<f4") == np.float32(1.0)))


@pytest.mark.skipif(not (os.path.exists(SENDER) and os.path.exists(SHIM)),
                    reason="oracle/_ref/ref_sender not built (needs /root/reference: make -C oracle ref)")
def test_reference_sender_speaks_the_format_the_ingest_expects():
    base = wire.free_base_port(THREAD_NUM)
    if base is None:
        pytest.skip("no free port range")
    total = 2 * TOTAL_BATCH_NUM     # sender.c:72: the shared counter runs to TOTAL_BATCH_NUM * 2
    srv = wire.CaptureServer(base, THREAD_NUM, BLOCK_SIZE, total, check=_all_ones)
    srv.start(timeout_s=240)
    env = dict(os.environ, LD_PRELOAD=SHIM, FR_SENDER_PORT_FROM=str(PORT), FR_SENDER_PORT_TO=str(base))
    p = subprocess.run([SENDER], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    seen = srv.join()
    out = p.stdout.decode(errors="replace")
    assert p.returncode == 0, out[-400:]
    # one connection per sender thread on PORT + i (sender.c:131-133), each a headerless stream of whole blocks of
    # BATCH_SIZE x INPUT_FEATURE_LEN little-endian fp32 (all 1.0: sender.c:32-35); the batch numbers come off ONE
    # counter shared by the connections, so only the total is fixed; then the sender waits for the server's word
    assert all(srv.accepted) and seen == total
    assert srv.stray == [0] * THREAD_NUM and srv.bad == [0] * THREAD_NUM
    assert all(b > 0 for b in srv.blocks)
    assert out.count("received from server: done") == THREAD_NUM


def test_python_sender_emulation_matches_the_capture_rules():
    """The emulated sender of the GPU ingest tests against the same capture server: whole blocks, no header, raw
    little-endian fp32 (or int32 index rows), any split over the connections."""
    n_conn, B, width = 3, 16, 48
    base = wire.free_base_port(n_conn)
    if base is None:
        pytest.skip("no free port range")
    per_conn = [2, 5, 1]
    srv = wire.CaptureServer(base, n_conn, B * width * 4, sum(per_conn), check=_all_ones)
    srv.start(timeout_s=60)
    ones = np.ones((B, width), np.float32)
    th = [threading.Thread(target=wire.send_blocks, args=(base + c, [ones] * per_conn[c])) for c in range(n_conn)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=60)
    assert srv.join() == sum(per_conn)
    assert srv.blocks == per_conn and srv.stray == [0] * n_conn and srv.bad == [0] * n_conn
