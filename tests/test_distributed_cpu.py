"""World-size-2 gloo test of the data-parallel wiring on the CPU: the reference-named
helpers (init_distributed_mode / reduce_value / ...), DDP over our custom autograd Functions, gradient ==
mean of per-shard gradients, parameters identical across ranks after a step, BatchNorm statistics local."""
import os
import socket
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo_data_parallel(tmp_path):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "_dist_worker.py"), str(r), "2", str(port), str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    for r in range(2):
        worst, same, bn_differs, flat_vs_ddp = open(tmp_path / f"rank{r}.txt").read().split()
        assert float(flat_vs_ddp) < 1e-6, "flat-buffer all-reduce step != DDP step"
        assert float(worst) < 1e-5, f"DDP gradient != mean of shard gradients ({worst})"
        assert same == "1", "parameters diverged across ranks after one step"
        assert bn_differs == "1", "BatchNorm statistics should stay per-process"
