"""`accelerate.Accelerator` when the package is installed, else a minimal stand-in with the members the
reference touches (src/train.py:61-64: `Accelerator()`, `.prepare(model, optimizer, lr_scheduler)`;
src/trainer/base.py:63,150: `.device`, `.backward(loss)`; `.is_main_process`, `.state`).

Launch contract is the one `accelerate launch` / `torchrun` set up: RANK, LOCAL_RANK, WORLD_SIZE,
MASTER_ADDR, MASTER_PORT.  One process per GPU; the process group (NCCL) is created when WORLD_SIZE > 1.
"""
from __future__ import annotations

import os

import torch

try:                                            # pragma: no cover - not installed in this image
    from accelerate import Accelerator          # noqa: F401
except Exception:
    class _State:
        def __init__(self, acc):
            self.num_processes = acc.num_processes
            self.process_index = acc.process_index
            self.local_process_index = acc.local_process_index
            self.device = acc.device
            self.distributed_type = "MULTI_GPU" if acc.num_processes > 1 else "NO"

    class Accelerator:
        def __init__(self, **_unused):
            import vsb200 as vs
            vs.require_b200()                    # no CPU path
            self.num_processes = int(os.environ.get("WORLD_SIZE", "1"))
            self.process_index = int(os.environ.get("RANK", "0"))
            self.local_process_index = int(os.environ.get("LOCAL_RANK", "0"))
            self.device = torch.device("cuda", self.local_process_index)
            torch.cuda.set_device(self.device)
            if self.num_processes > 1 and not torch.distributed.is_initialized():
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                torch.distributed.init_process_group(backend="nccl", rank=self.process_index, world_size=self.num_processes)
            self.state = _State(self)

        @property
        def is_main_process(self):
            return self.process_index == 0

        @property
        def is_local_main_process(self):
            return self.local_process_index == 0

        def prepare(self, *objs):
            out = []
            for o in objs:
                if isinstance(o, torch.nn.Module):
                    o = o.to(self.device)
                out.append(o)
            return out[0] if len(out) == 1 else tuple(out)

        def backward(self, loss, **kw):
            loss.backward(**kw)

        def wait_for_everyone(self):
            if self.num_processes > 1:
                torch.distributed.barrier()

        def print(self, *a, **k):
            if self.is_main_process:
                print(*a, **k)
