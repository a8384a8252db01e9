"""npz-backed stand-in for the tiny part of h5py the reference's fluid_reference uses (test tooling only)."""
import os
import numpy as np
class _DS:
    def __init__(self, arr): self._a = np.asarray(arr)
    def __getitem__(self, k): return self._a[k]
    def __array__(self, dtype=None, copy=None): return self._a if dtype is None else self._a.astype(dtype)
    @property
    def shape(self): return self._a.shape
class File:
    def __init__(self, path, mode="r"):
        self.path, self.mode, self._d = path, mode, {}
        if mode == "r":
            if not os.path.exists(path): raise OSError(path)
            with np.load(path + ".npz" if not path.endswith(".npz") and os.path.exists(path + ".npz") else path, allow_pickle=False) as z:
                self._d = {k: z[k] for k in z.files}
    def create_dataset(self, name, data=None): self._d[name] = np.asarray(data)
    def __getitem__(self, name): return _DS(self._d[name])
    def close(self):
        if self.mode == "w":
            with open(self.path, "wb") as f: np.savez(f, **self._d)
    def __enter__(self): return self
    def __exit__(self, *a): self.close()
