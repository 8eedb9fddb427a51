"""A deliberately tiny column container standing in for cudf.DataFrame on this path.

The reference returns cudf objects; requiring cuDF/RMM on the hot path is exactly what this
package avoids, so results are ordered {column name -> device tensor} maps with the reference's
column names and dtypes.
"""
from collections import OrderedDict


class Frame:
    def __init__(self, columns):
        self._cols = OrderedDict(columns)

    @property
    def columns(self):
        return list(self._cols.keys())

    def __getitem__(self, name):
        return self._cols[name]

    def __contains__(self, name):
        return name in self._cols

    def __len__(self):
        for v in self._cols.values():
            return int(v.shape[0])
        return 0

    def __iter__(self):
        return iter(self._cols)

    def items(self):
        return self._cols.items()

    @property
    def dtypes(self):
        return {k: v.dtype for k, v in self._cols.items()}

    def to_numpy(self):
        """Host copies with the exact dtypes (uint32 / uint8 / bool)."""
        return {k: v.cpu().numpy() for k, v in self._cols.items()}

    def to_pandas(self):
        import pandas as pd

        return pd.DataFrame(self.to_numpy())

    def __repr__(self):
        return "Frame(%s, rows=%d)" % (", ".join("%s:%s" % (k, str(v.dtype).replace("torch.", ""))
                                                 for k, v in self._cols.items()), len(self))
