"""Drive the UNMODIFIED reference BWAS binary (oracle/_ref/parallel_weighted_astar, compiled from
/root/reference/cpp/*.cpp by oracle/Makefile) the way search_methods/astar.py:457-568 does, with the
heuristic served over the AF_UNIX socket protocol of astar.py:571-616 / parallel_weighted_astar.cpp:121-136,
275-279.  TEST INFRASTRUCTURE ONLY (also used by bench.py --impl reference / cpu_baseline).

Wire protocol (restated): request = uint64 nbytes + uint8[n*S] child states; reply = float32[n].
"""
from __future__ import annotations

import os
import socket
import subprocess
import tempfile
import threading
from typing import Callable, Dict, List, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BINARY = os.path.join(HERE, "_ref", "parallel_weighted_astar")


def have_reference_binary() -> bool:
    return os.path.exists(REF_BINARY) and os.access(REF_BINARY, os.X_OK)


def _recv_exact(conn: socket.socket, n: int) -> bytes:
    chunks = []
    got = 0
    while got < n:
        c = conn.recv(min(1 << 20, n - got))
        if not c:
            raise ConnectionError("peer closed")
        chunks.append(c); got += len(c)
    return b"".join(chunks)


class HeuristicServer:
    """cpp_listener (astar.py:571-616): one connection at a time, re-accept when the client goes away."""

    def __init__(self, state_dim: int, heuristic: Callable[[np.ndarray], np.ndarray], socket_path: Optional[str] = None):
        self.state_dim = state_dim
        self.heuristic = heuristic
        self.dir = tempfile.mkdtemp(prefix="dcb_sock_")
        self.path = socket_path or os.path.join(self.dir, "h.sock")
        self.sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self.sock.bind(self.path)
        self.sock.listen(1)
        self.num_states_served = 0
        self.heur_seconds = 0.0
        self._stop = False
        self.thread = threading.Thread(target=self._serve, daemon=True)
        self.thread.start()

    def _serve(self):
        import time
        self.sock.settimeout(0.2)
        while not self._stop:
            try:
                conn, _ = self.sock.accept()
            except socket.timeout:
                continue
            except OSError:
                return
            conn.settimeout(None)
            try:
                while True:
                    hdr = conn.recv(8)
                    if not hdr:
                        break
                    if len(hdr) < 8:
                        hdr += _recv_exact(conn, 8 - len(hdr))
                    nbytes = int(np.frombuffer(hdr, dtype=np.int64)[0])
                    data = _recv_exact(conn, nbytes)
                    st = np.frombuffer(data, dtype=np.uint8).reshape(-1, self.state_dim)
                    t0 = time.time()
                    h = np.asarray(self.heuristic(st), dtype=np.float32)
                    self.heur_seconds += time.time() - t0
                    self.num_states_served += st.shape[0]
                    conn.sendall(np.ascontiguousarray(np.maximum(h, np.float32(0.0))).tobytes())   # clip_zero=True (astar.py:491-493)
            except (ConnectionError, OSError):
                pass
            finally:
                conn.close()

    def close(self):
        self._stop = True
        try:
            self.sock.close()
        finally:
            try:
                os.unlink(self.path)
                os.rmdir(self.dir)
            except OSError:
                pass


def run_reference_bwas(env_name: str, state: np.ndarray, weight: float, batch_size: int, server: HeuristicServer,
                       timeout: Optional[float] = None, omp_threads: Optional[int] = None) -> Dict:
    """astar.py:508-532: spawn the binary for ONE start state, parse moves / nodes generated / total time."""
    if not have_reference_binary():
        raise RuntimeError("oracle/_ref/parallel_weighted_astar is not built (make -C oracle ref)")
    state_str = " ".join(str(int(x)) for x in state)
    env = dict(os.environ)
    if omp_threads:
        env["OMP_NUM_THREADS"] = str(omp_threads)
    out = subprocess.run([REF_BINARY, state_str, str(weight), str(batch_size), server.path, env_name.lower(), "0"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, env=env)
    lines = out.stdout.split("\n")
    if lines and lines[-1] == "":
        lines = lines[:-1]
    moves = [int(x) for x in lines[-5].split(" ")[:-1]][::-1]        # printed goal->root (:336-341), reversed (astar.py:530)
    phase = {"exp": 0.0, "check": 0.0, "write": 0.0, "heur": 0.0, "remOpen": 0.0, "add": 0.0, "cost": 0.0}
    iters = 0
    for ln in lines:
        if ln.startswith("Times - "):
            iters += 1
            for part in ln[len("Times - "):].split(", "):
                k, _, v = part.partition(": ")
                if k in phase:
                    phase[k] += float(v)
    return {"moves": moves, "nodes_generated": int(lines[-3]), "time": float(lines[-1]), "iterations": iters, "phase_seconds": phase}
