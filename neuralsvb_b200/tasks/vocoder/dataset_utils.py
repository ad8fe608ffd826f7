"""Dataset-backed input of the vocoder task: the reference's IndexedDataset on disk
(utils/indexed_datasets.py:7-54; items written by the binarizer: ``mel [T, 80]``, ``wav`` float16 ``[T*hop]``, ``f0 [T]``)
-> pinned, batched host tensors -> one non-blocking H2D per batch.

The reference path is pickle-per-item reads in DataLoader workers, a Python collate with per-sample copies
(utils/__init__.py:118-150) and ``pin_memory=False`` (tasks/tts/tts.py:97-101).  Here a producer thread reads and
un-pickles items, crops ``max_samples`` (a multiple of hop, frame aligned: ``wav[s*hop:(s+T)*hop]`` with ``mel[s:s+T]``
and ``f0[s:s+T]``) and writes them straight into a ring of page-locked batch buffers, so the trainer's ``move_to_cuda``
is a single async copy per tensor that overlaps the previous step.  Batches are built over the GLOBAL batch
(``max_sentences`` x world size) and rank r takes ``batch[r::world]`` (tasks/tts/tts.py:59-72,93-96)."""
import pickle
import queue
import threading

import numpy as np
import torch

from neuralsvb_b200.utils.indexed_datasets import IndexedDataset


class VocoderBatchLoader:
    def __init__(self, path, hop, max_samples, max_sentences, rank=0, world=1, seed=1234, shuffle=True, endless=False,
                 n_mel=80, depth=3, pin=None, device=None):
        self.path, self.hop, self.n_mel = path, int(hop), int(n_mel)
        self.T = int(max_samples) // self.hop
        self.B, self.rank, self.world = int(max_sentences), int(rank), int(world)
        self.seed, self.shuffle, self.endless, self.depth = int(seed), shuffle, endless, int(depth)
        self.lengths = None
        self.pin = torch.cuda.is_available() if pin is None else pin
        # device given: the loader owns the H2D too (copy stream + event; the consumer's stream waits on the event and
        # the page-locked buffer returns to the ring only when its copy has completed)
        self.device = None if device is None else torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device) if self.device is not None else None
        n = len(IndexedDataset(path))
        assert n > 0, f'{path}: empty dataset'
        self.n_items = n
        self.epoch = 0

    def __len__(self):
        return self.n_items // (self.B * self.world)

    def _buffers(self):
        mk = lambda *shape: torch.empty(*shape, dtype=torch.float32).pin_memory() if self.pin else torch.empty(*shape, dtype=torch.float32)
        return {'wavs': mk(self.B, 1, self.T * self.hop), 'mels': mk(self.B, self.T, self.n_mel), 'f0': mk(self.B, self.T)}

    def _global_batches(self, epoch):
        rs = np.random.RandomState(self.seed + epoch)
        idx = rs.permutation(self.n_items) if self.shuffle else np.arange(self.n_items)
        gb = self.B * self.world
        for i in range(0, len(idx) - gb + 1, gb):              # batches that do not divide evenly are dropped (tts.py:69-72)
            yield epoch, i // gb, idx[i:i + gb][self.rank::self.world]

    def _fill(self, ds, buf, epoch, bi, ids):
        rs = np.random.RandomState((self.seed * 1000003 + epoch * 10007 + bi) % (2 ** 31) + self.rank)
        names = []
        for j, i in enumerate(ids):
            item = pickle.loads(ds.read_raw(int(i)))
            mel, wav = item['mel'], item['wav']
            frames = min(len(mel), len(wav) // self.hop)
            T = self.T
            s = int(rs.randint(0, frames - T + 1)) if frames > T else 0
            t = min(T, frames)
            buf['mels'][j].zero_(), buf['wavs'][j].zero_(), buf['f0'][j].zero_()
            buf['mels'][j, :t] = torch.from_numpy(np.ascontiguousarray(mel[s:s + t], dtype=np.float32))
            buf['wavs'][j, 0, :t * self.hop] = torch.from_numpy(wav[s * self.hop:(s + t) * self.hop].astype(np.float32))
            if item.get('f0') is not None:
                buf['f0'][j, :t] = torch.from_numpy(np.asarray(item['f0'][s:s + t], dtype=np.float32))
            names.append(item.get('item_name'))
        return names

    def __iter__(self):
        ring = [self._buffers() for _ in range(self.depth)]
        free, ready = queue.Queue(), queue.Queue(maxsize=self.depth)
        for b in ring:
            free.put(b)
        stop = threading.Event()

        def produce():
            ds = IndexedDataset(self.path, num_cache=0)        # own file handle: seeks do not race the caller's
            try:
                epoch = self.epoch
                while not stop.is_set():
                    for ep, bi, ids in self._global_batches(epoch):
                        buf = free.get()
                        if stop.is_set():
                            return
                        names = self._fill(ds, buf, ep, bi, ids)
                        ready.put((buf, names))
                    epoch += 1
                    if not self.endless:
                        break
                ready.put(None)
            except Exception as e:                             # surface loader errors in the consumer
                ready.put(e)
        th = threading.Thread(target=produce, daemon=True)
        th.start()
        prev, pending = None, []
        try:
            while True:
                while pending and (pending[0][0].query() or ready.empty()):
                    ev, b = pending.pop(0)
                    ev.synchronize()
                    free.put(b)                                # its H2D copy has completed
                got = ready.get()
                if prev is not None:
                    free.put(prev)                             # host consumers: the batch handed out last time has been consumed
                    prev = None
                if got is None:
                    break
                if isinstance(got, Exception):
                    raise got
                buf, names = got
                if self.device is None:
                    prev = buf
                    yield {'wavs': buf['wavs'], 'mels': buf['mels'], 'f0': buf['f0'], 'item_names': names}
                    continue
                cur = torch.cuda.current_stream(self.device)
                with torch.cuda.stream(self.copy_stream):
                    out = {k: buf[k].to(self.device, non_blocking=True) for k in ('wavs', 'mels', 'f0')}
                    ev = torch.cuda.Event()
                    ev.record(self.copy_stream)
                cur.wait_event(ev)
                for t in out.values():
                    t.record_stream(cur)
                pending.append((ev, buf))
                out['item_names'] = names
                yield out
        finally:
            stop.set()
            try:
                free.put_nowait(ring[0])
            except Exception:
                pass
            self.epoch += 1
