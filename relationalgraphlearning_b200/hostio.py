"""Host-buffer front end: stream batches that live in (pinned) host memory through the CUDA path.

The reference's callers hand CPU tensors to the modules (crowd_nav/utils/trainer.py:75-80,
crowd_sim/envs/utils/state.py:64-79).  `HostStream` is the equivalent entry point here: each
`submit()` enqueues  host->device copy of the states, the kernels, and the device->host copy of the
result on three CUDA streams over a ring of device/pinned buffers, so consecutive batches overlap
(PCIe in, SMs, PCIe out) and the host never blocks until it asks for a result.
"""
import torch


class HostStream(object):
    KINDS = ('graph', 'value', 'statepred')

    def __init__(self, kind, module, batch, human_num, device, depth=4):
        assert kind in self.KINDS
        self.kind, self.module, self.B, self.Nh, self.dev, self.depth = kind, module, batch, human_num, device, depth
        n = human_num + 1
        out_shape = {'graph': (batch, n, 32), 'value': (batch, 1), 'statepred': (batch, human_num, 5)}[kind]
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(device) for _ in range(3))
        self.robot_d = [torch.empty(batch, 1, 9, device=device) for _ in range(depth)]
        self.humans_d = [torch.empty(batch, human_num, 5, device=device) for _ in range(depth)]
        self.out_h = [torch.empty(out_shape, pin_memory=True) for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_run = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.ev_free = [None] * depth          # device input slot consumed by the kernels
        self.count = 0
        self.h2d_bytes = batch * (9 + 5 * human_num) * 4
        self.d2h_bytes = int(torch.tensor(out_shape).prod()) * 4

    def _run(self, robot, humans):
        if self.kind == 'graph':
            return self.module.run(robot, humans, want_H=True)['H']
        if self.kind == 'value':
            return self.module.run(robot, humans)
        return self.module.run(robot, humans)

    def submit(self, robot_h, humans_h):
        """robot_h[B,1,9], humans_h[B,Nh,5]: CPU tensors (pinned for truly asynchronous copies).  Returns slot."""
        k = self.count % self.depth
        self.count += 1
        with torch.cuda.stream(self.s_in):
            if self.ev_free[k] is not None:
                self.s_in.wait_event(self.ev_free[k])      # kernels of the previous user of this slot are done
            self.robot_d[k].copy_(robot_h, non_blocking=True)
            self.humans_d[k].copy_(humans_h, non_blocking=True)
            self.ev_in[k].record(self.s_in)
        with torch.cuda.stream(self.s_run), torch.no_grad():
            self.s_run.wait_event(self.ev_in[k])
            out = self._run(self.robot_d[k], self.humans_d[k])
            self.ev_run[k].record(self.s_run)
            self.ev_free[k] = self.ev_run[k]
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_run[k])
            self.out_h[k].copy_(out, non_blocking=True)
            out.record_stream(self.s_out)
            self.ev_out[k].record(self.s_out)
        return k

    def result(self, slot):
        """Blocks until the slot's result has landed in pinned host memory and returns it."""
        self.ev_out[slot].synchronize()
        return self.out_h[slot]

    def drain(self):
        self.s_out.synchronize()
