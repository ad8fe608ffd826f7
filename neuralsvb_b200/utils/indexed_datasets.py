"""The reference's on-disk dataset format (utils/indexed_datasets.py:7-54), kept byte for byte: ``{path}.data`` =
concatenated pickles of the items, ``{path}.idx`` = ``np.save`` of ``{'offsets': [...]}``.  A dataset written by the
reference's binarizer opens here and vice versa."""
import pickle

import numpy as np


class IndexedDataset:
    def __init__(self, path, num_cache=1):
        self.path = path
        self.data_file = None
        self.data_offsets = np.load(f'{path}.idx', allow_pickle=True).item()['offsets']
        self.data_file = open(f'{path}.data', 'rb', buffering=-1)
        self.cache, self.num_cache = [], num_cache

    def check_index(self, i):
        if i < 0 or i >= len(self.data_offsets) - 1:
            raise IndexError('index out of range')

    def __del__(self):
        if self.data_file:
            self.data_file.close()

    def read_raw(self, i):
        """The pickled bytes of item i (one seek + one read; the loader un-pickles in its worker thread)."""
        self.check_index(i)
        self.data_file.seek(self.data_offsets[i])
        return self.data_file.read(self.data_offsets[i + 1] - self.data_offsets[i])

    def __getitem__(self, i):
        self.check_index(i)
        for c in self.cache:
            if c[0] == i:
                return c[1]
        item = pickle.loads(self.read_raw(i))
        if self.num_cache > 0:
            self.cache = [(i, item)] + self.cache[:self.num_cache - 1]
        return item

    def __len__(self):
        return len(self.data_offsets) - 1


class IndexedDatasetBuilder:
    def __init__(self, path):
        self.path = path
        self.out_file = open(f'{path}.data', 'wb')
        self.byte_offsets = [0]

    def add_item(self, item):
        n = self.out_file.write(pickle.dumps(item))
        self.byte_offsets.append(self.byte_offsets[-1] + n)

    def finalize(self):
        self.out_file.close()
        with open(f'{self.path}.idx', 'wb') as f:
            np.save(f, {'offsets': self.byte_offsets})
