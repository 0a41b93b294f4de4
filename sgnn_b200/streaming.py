"""Host-buffer inference pipeline: scene batches arrive in (pinned) host memory and results are wanted back on the
host -- the way test_scene.py consumes the model (inputs from the DataLoader, `.cpu().numpy()` on the outputs,
test_scene.py:81,98).  `StreamingRunner` overlaps the three phases on three CUDA streams:

    copy-in stream :  H2D of batch i+1        (pinned -> device staging, double buffered)
    compute stream :  GenModel.forward(batch i)
    copy-out stream:  D2H of result i-1       (device -> pinned)

so the PCIe transfers (15 MB in, 24 MB out per 32-block step) hide behind the ~7.5 ms of compute.  Results are
identical to calling the model directly.
"""
import collections

import torch


class StreamingRunner(object):
    def __init__(self, model, depth=2, overlap_in=True, overlap_out=True):
        self.model = model
        self.dev = next(model.parameters()).device
        self.depth = depth
        cur = torch.cuda.current_stream(self.dev)
        self.s_in = torch.cuda.Stream(device=self.dev) if overlap_in else cur
        self.s_out = torch.cuda.Stream(device=self.dev) if overlap_out else cur
        self.slots = [dict(locs=None, feats=None, ev=None) for _ in range(depth)]
        self.pins = [dict(locs=None, sdf=None) for _ in range(depth)]
        self.pending = collections.deque()     # submitted, H2D enqueued, not yet computed
        self.n_sub = 0

    def _stage(self, slot, host_locs, host_feats):
        """Enqueue the H2D copies of one batch on the copy-in stream."""
        s = self.slots[slot]
        if s['locs'] is None or s['locs'].shape[0] < host_locs.shape[0]:
            n = int(host_locs.shape[0] * 1.25) + 1
            s['locs'] = torch.empty((n, 4), dtype=host_locs.dtype, device=self.dev)
            s['feats'] = torch.empty((n, host_feats.shape[1]), dtype=torch.float32, device=self.dev)
        n = host_locs.shape[0]
        with torch.cuda.stream(self.s_in):
            # the staging slot was last read by the forward `depth` submissions ago: wait for that compute
            if s['ev'] is not None:
                self.s_in.wait_event(s['ev'])
            s['locs'][:n].copy_(host_locs, non_blocking=True)
            s['feats'][:n].copy_(host_feats, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.s_in)
        return n, ready

    def submit(self, host_locs, host_feats, batch_size):
        """Queue one batch (CPU tensors, ideally pinned).  Returns nothing; call `step()` to run the oldest queued
        batch and obtain a ticket for its host-side result."""
        slot = self.n_sub % self.depth
        n, ready = self._stage(slot, host_locs, host_feats)
        self.pending.append((slot, n, ready, int(batch_size)))
        self.n_sub += 1

    def step(self, loss_weights):
        """Run the oldest submitted batch on the current stream; enqueue the D2H of its result on the copy-out stream.
        Returns a ticket for `result()`."""
        slot, n, ready, bs = self.pending.popleft()
        s = self.slots[slot]
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(ready)
        (ol, osdf), levels = self.model([s['locs'][:n], s['feats'][:n], bs], loss_weights)
        done = torch.cuda.Event()
        done.record(cur)
        s['ev'] = done
        m = 0 if isinstance(ol, list) else int(ol.shape[0])
        p = self.pins[slot]
        if m and (p['locs'] is None or p['locs'].shape[0] < m):
            cap = int(m * 1.25) + 1
            p['locs'] = torch.empty((cap, 4), dtype=ol.dtype).pin_memory()
            p['sdf'] = torch.empty((cap, osdf.shape[1]), dtype=torch.float32).pin_memory()
        out_ev = torch.cuda.Event()
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            if m:
                p['locs'][:m].copy_(ol, non_blocking=True)
                p['sdf'][:m].copy_(osdf, non_blocking=True)
                ol.record_stream(self.s_out)       # the caching allocator may recycle them once the copies are done
                osdf.record_stream(self.s_out)
            out_ev.record(self.s_out)
        # NOTE for callers: `levels` (per-hierarchy device outputs) and every live ticket pin device memory; holding
        # many of them forces the allocator into fresh cudaMallocs each step.  Drop tickets once consumed.
        return dict(slot=slot, m=m, ev=out_ev, levels=levels)

    def result(self, ticket):
        """Host tensors (views of the pinned result buffers, valid until `depth` more steps) of a finished step."""
        ticket['ev'].synchronize()
        p = self.pins[ticket['slot']]
        m = ticket['m']
        if m == 0:
            return [], []
        return p['locs'][:m], p['sdf'][:m]

    def drain(self):
        torch.cuda.current_stream(self.dev).wait_stream(self.s_out)
