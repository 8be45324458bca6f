#!/usr/bin/env python3
"""oracle/mpirun_shim.py -- TEST INFRASTRUCTURE, not product code.

`mpirun -np P <binary> args...` stand-in for binaries compiled against
oracle/mpi_shim/mpi.h: creates the per-pair FIFOs, starts P processes with
MPISHIM_RANK/SIZE/DIR set, waits for all of them and returns the worst exit
code.  Rank r's stdout/stderr go to <logdir>/rank<r>.{out,err} when --logdir
is given (rank 0 is echoed otherwise).
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile


def run(np_, cmd, logdir=None, timeout=None, env_extra=None):
    d = tempfile.mkdtemp(prefix="mpishim_")
    procs = []
    files = []
    try:
        for a in range(np_):
            for b in range(np_):
                if a != b:
                    os.mkfifo(os.path.join(d, f"f_{a}_{b}"))
        for r in range(np_):
            env = dict(os.environ)
            env.update({"MPISHIM_RANK": str(r), "MPISHIM_SIZE": str(np_), "MPISHIM_DIR": d})
            if env_extra:
                env.update(env_extra)
            if logdir:
                os.makedirs(logdir, exist_ok=True)
                fo = open(os.path.join(logdir, f"rank{r}.out"), "wb")
                fe = open(os.path.join(logdir, f"rank{r}.err"), "wb")
                files += [fo, fe]
            else:
                fo = None if r == 0 else subprocess.DEVNULL
                fe = None if r == 0 else subprocess.DEVNULL
            procs.append(subprocess.Popen(cmd, env=env, stdout=fo, stderr=fe, stdin=subprocess.DEVNULL))
        rc = 0
        for p in procs:
            try:
                rc = max(rc, abs(p.wait(timeout=timeout)))
            except subprocess.TimeoutExpired:
                for q in procs:
                    if q.poll() is None:
                        q.kill()
                rc = 124
        return rc
    finally:
        for f in files:
            f.close()
        shutil.rmtree(d, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-np", type=int, required=True)
    ap.add_argument("--logdir", default=None)
    ap.add_argument("--timeout", type=float, default=None)
    ap.add_argument("cmd", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    cmd = a.cmd[1:] if a.cmd and a.cmd[0] == "--" else a.cmd
    sys.exit(run(a.np, cmd, a.logdir, a.timeout))


if __name__ == "__main__":
    main()
