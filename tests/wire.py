"""The reference's B2 wire format as test plumbing: a sender that behaves like
GPU/final_network_cublasLt_1_node_no_FIFO_scatter/multiple_connections_network_client_sender.c:55-100 (connect, then
send() raw little-endian bytes block after block, no header, no framing) and a capture server that receives the way
cuda_server.c:425-440 does (whole BLOCK_SIZE blocks per connection).  tests/test_wire_format.py runs the reference's own
sender binary against the capture server and checks that this emulation produces the same stream; the GPU ingest tests
drive fr_ingest_* with the emulation."""
import select
import socket
import threading
import time

import numpy as np


def send_blocks(port, blocks, host="127.0.0.1"):
    with socket.create_connection((host, port), timeout=30) as s:
        for blk in blocks:
            s.sendall(blk.tobytes())


def free_base_port(n, lo=18080, hi=28080):
    """First base with n consecutive free loopback ports, or None."""
    for base in range(lo, hi, 97):
        socks = []
        try:
            for i in range(n):
                s = socket.socket()
                s.bind(("127.0.0.1", base + i))
                socks.append(s)
            return base
        except OSError:
            continue
        finally:
            for s in socks:
                s.close()
    return None


class CaptureServer:
    """n_conn listening ports; every connection is read in whole blocks of block_bytes until `total_blocks` blocks have
    arrived over all connections together (the senders share one batch counter), then every sender gets `reply`."""

    def __init__(self, base_port, n_conn, block_bytes, total_blocks, reply=b"done", check=None):
        self.block_bytes, self.total_blocks, self.reply, self.check = block_bytes, total_blocks, reply, check
        self.blocks = [0] * n_conn            # whole blocks per connection
        self.stray = [0] * n_conn             # bytes of an unfinished block when the run ended
        self.bad = [0] * n_conn               # blocks the check refused
        self.accepted = [False] * n_conn
        self._lock = threading.Lock()
        self._seen = 0
        self._deadline = None
        self._lsn = []
        for i in range(n_conn):
            s = socket.socket()
            s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            s.bind(("127.0.0.1", base_port + i))
            s.listen(1)
            self._lsn.append(s)
        self._threads = [threading.Thread(target=self._serve, args=(i,), daemon=True) for i in range(n_conn)]

    def start(self, timeout_s):
        self._deadline = time.monotonic() + timeout_s
        for t in self._threads:
            t.start()

    def _done(self):
        with self._lock:
            return self._seen >= self.total_blocks

    def _serve(self, i):
        lsn = self._lsn[i]
        lsn.settimeout(max(self._deadline - time.monotonic(), 0.1))
        try:
            conn, _ = lsn.accept()
        except OSError:
            return
        self.accepted[i] = True
        buf = bytearray(self.block_bytes)
        view = memoryview(buf)
        arr = np.frombuffer(buf, dtype=np.uint8)
        got = 0
        with conn:
            conn.setblocking(False)
            while time.monotonic() < self._deadline:
                if got == 0 and self._done():
                    break
                r, _, _ = select.select([conn], [], [], 0.05)
                if not r:
                    continue
                try:
                    n = conn.recv_into(view[got:], self.block_bytes - got)
                except BlockingIOError:
                    continue
                if n == 0:                      # sender closed
                    break
                got += n
                if got == self.block_bytes:
                    if self.check is not None and not self.check(arr):
                        self.bad[i] += 1
                    self.blocks[i] += 1
                    got = 0
                    with self._lock:
                        self._seen += 1
            self.stray[i] = got
            # the reference sender blocks in read() until the server answers (sender.c:121-123)
            while not self._done() and time.monotonic() < self._deadline:
                time.sleep(0.01)
            try:
                conn.setblocking(True)
                conn.sendall(self.reply)
            except OSError:
                pass

    def join(self):
        for t in self._threads:
            t.join(timeout=max(self._deadline - time.monotonic(), 0) + 5)
        for s in self._lsn:
            s.close()
        return sum(self.blocks)
