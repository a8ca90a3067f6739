"""GPU-resident replay memory -- drop-in for crowd_nav/utils/memory.py:4-28 (`ReplayMemory`) on the MPRL path.

The reference keeps a python list of 6-tuples of tiny tensors and lets torch's DataLoader collate `batch_size` of them per
minibatch (crowd_nav/utils/trainer.py:66-67,113-114): for the 8 192-sample step of BASELINE configs[3] that is ~50 000
python-level tensor operations per minibatch.  `DeviceReplayMemory` stores every transition as one contiguous record in
HBM and builds a minibatch with ONE gather launch (csrc/replay.cu).

Same surface as the reference class (`push`, `is_full`, `__getitem__`, `__len__`, `clear`, `capacity`, `position`) so
`Explorer.update_memory` (crowd_nav/utils/explorer.py:113-140) can push into it unchanged, plus
  * `sample(batch_size)`                one random minibatch (indices drawn on the device),
  * `loader(batch_size, shuffle=True)`  an iterable with the DataLoader protocol the reference trainer uses
                                        (`for data in loader`, `len(loader)`): assign it to `MPRLTrainer.data_loader`
                                        (the trainer only builds its own DataLoader when that attribute is None).
The epoch order of `loader` is a random permutation drawn from torch's default CPU generator, exactly like
DataLoader(shuffle=True)'s RandomSampler draws its seed, so a seeded run visits the same batches.
"""
import torch

from . import _lib, ops


class DeviceReplayMemory(object):
    def __init__(self, capacity, human_num, device):
        self.capacity = int(capacity)
        self.human_num = int(human_num)
        self.device = torch.device(device)
        self.rec = int(_lib.lib().rgl_replay_record_floats(self.human_num))
        if self.rec == 0:
            raise _lib.RglError('DeviceReplayMemory: human count outside [1,31]')
        self.store = torch.zeros(self.capacity, self.rec, dtype=torch.float32, device=self.device)
        self.size = 0
        self.position = 0

    # ---- crowd_nav/utils/memory.py surface ----
    def push(self, item):
        """item = (robot[1,9], humans[Nh,5], value[1], reward[1], next_robot[1,9], next_humans[Nh,5]) as Explorer pushes them."""
        t = [ops._f32c(torch.as_tensor(x).to(self.device)) for x in item]
        assert t[0].numel() == 9 and t[4].numel() == 9 and t[1].numel() == 5 * self.human_num and t[5].numel() == 5 * self.human_num
        with torch.cuda.device(self.device):
            rc = _lib.lib().rgl_replay_push(_lib.ptr(self.store), self.position, self.human_num, *[_lib.ptr(x) for x in t],
                                            _lib.stream_ptr(self.device))
        _lib.check(rc, 'rgl_replay_push')
        ops._count(1)
        self.size = max(self.size, self.position + 1)
        self.position = (self.position + 1) % self.capacity

    def push_batch(self, robot, humans, value, reward, next_robot, next_humans):
        """Many transitions at once (vectorised environments): tensors with a leading batch dimension."""
        n = robot.size(0)
        rows = torch.cat([robot.reshape(n, -1), humans.reshape(n, -1), value.reshape(n, 1), reward.reshape(n, 1),
                          next_robot.reshape(n, -1), next_humans.reshape(n, -1)], dim=1).to(self.device, torch.float32)
        assert rows.size(1) == self.rec
        slots = (torch.arange(n, device=self.device) + self.position) % self.capacity
        self.store.index_copy_(0, slots, rows)
        self.size = min(self.capacity, max(self.size, self.position + n))
        self.position = (self.position + n) % self.capacity

    def is_full(self):
        return self.size == self.capacity

    def __len__(self):
        return self.size

    def __getitem__(self, i):
        if not -self.size <= i < self.size:
            raise IndexError(i)
        r, Nh = self.store[i], self.human_num
        hw = 5 * Nh
        return (r[0:9].view(1, 9), r[9:9 + hw].view(Nh, 5), r[9 + hw:10 + hw], r[10 + hw:11 + hw],
                r[11 + hw:20 + hw].view(1, 9), r[20 + hw:20 + 2 * hw].view(Nh, 5))

    def clear(self):
        self.size = 0
        self.position = 0

    # ---- minibatches ----
    def gather(self, idx):
        """idx: int64 device tensor [B] -> the trainer's 6-tuple (robot[B,1,9], humans[B,Nh,5], value[B,1], reward[B,1],
        next_robot[B,1,9], next_humans[B,Nh,5]) in one launch."""
        idx = idx.to(self.device, torch.int64).contiguous()
        B, Nh, dev = idx.numel(), self.human_num, self.device
        out = (torch.empty(B, 1, 9, device=dev), torch.empty(B, Nh, 5, device=dev), torch.empty(B, 1, device=dev),
               torch.empty(B, 1, device=dev), torch.empty(B, 1, 9, device=dev), torch.empty(B, Nh, 5, device=dev))
        with torch.cuda.device(dev):
            rc = _lib.lib().rgl_replay_gather(_lib.ptr(self.store), _lib.ptr(idx), B, Nh, *[_lib.ptr(x) for x in out], _lib.stream_ptr(dev))
        _lib.check(rc, 'rgl_replay_gather')
        ops._count(1 if B else 0)
        return out

    def sample(self, batch_size):
        return self.gather(torch.randint(0, self.size, (batch_size,), device=self.device))

    def loader(self, batch_size, shuffle=True):
        return _Loader(self, batch_size, shuffle)


class _Loader(object):
    """DataLoader-protocol view (`for data in loader`, `len(loader)`) over a DeviceReplayMemory."""

    def __init__(self, memory, batch_size, shuffle):
        self.memory, self.batch_size, self.shuffle = memory, int(batch_size), shuffle

    def __len__(self):
        return (len(self.memory) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.memory)
        if self.shuffle:
            # the same draws DataLoader(shuffle=True) makes from the default generator: the iterator's base seed first,
            # then RandomSampler seeds a fresh generator for the permutation (so a seeded run visits the same batches)
            torch.empty((), dtype=torch.int64).random_()
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            g = torch.Generator()
            g.manual_seed(seed)
            order = torch.randperm(n, generator=g)
        else:
            order = torch.arange(n)
        order = order.to(self.memory.device)
        for lo in range(0, n, self.batch_size):
            yield self.memory.gather(order[lo:lo + self.batch_size])
