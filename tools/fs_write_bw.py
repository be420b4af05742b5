"""Ceiling of the output side of a file-level run on this box: N threads appending 64 MB blocks to N files in `dir`."""
import os
import sys
import tempfile
import threading
import time

d = sys.argv[1] if len(sys.argv) > 1 else tempfile.gettempdir()
block = os.urandom(1 << 20) * 64
for n in (1, 4, 16):
    paths = [os.path.join(d, "qcb_fsbw_%d" % i) for i in range(n)]

    def work(p):
        fd = os.open(p, os.O_WRONLY | os.O_CREAT | os.O_TRUNC | os.O_APPEND, 0o666)
        for _ in range(8):
            os.write(fd, block)
        os.close(fd)
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(p,)) for p in paths]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    print("%2d writers: %.2f GB/s (%s)" % (n, n * 8 * len(block) / 1e9 / dt, d))
    for p in paths:
        os.remove(p)
