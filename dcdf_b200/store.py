"""Stores for the stored nodes: anything that maps a CID (36 bytes) to the node's bytes works as `store` for Variable /
Dataset / Superchunk.open -- the role of the reference's `Mapper` (mapper.rs:10-38: `store()` hands out a writer whose
`finish()` returns the CID, `load(cid)` a reader).  A dict is the reference's in-memory test mapper (testing.rs:101-183);
`DirStore` keeps one file per object so that a dataset survives the process, as dcdf-ipfs does with a local IPFS node.
"""
import os
import tempfile

from .span import CID_BYTES, cid_of


class DirStore:
    """One file per stored object, named by the hex CID, under `root` (two-level fan-out).  Objects are immutable and
    content addressed: writing an existing CID is a no-op, a write is atomic (temp file + rename)."""

    def __init__(self, root, verify=False):
        self.root, self.verify = os.fspath(root), bool(verify)
        os.makedirs(self.root, exist_ok=True)

    def _path(self, cid):
        cid = bytes(cid)
        if len(cid) != CID_BYTES:
            raise KeyError(cid)
        h = cid.hex()
        return os.path.join(self.root, h[8:10], h)

    def __contains__(self, cid):
        try:
            return os.path.exists(self._path(cid))
        except KeyError:
            return False

    def __setitem__(self, cid, data):
        path = self._path(cid)
        if os.path.exists(path):
            return
        data = bytes(data)
        if self.verify and cid_of(data) != bytes(cid):
            raise ValueError("object does not hash to its CID")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        fd, tmp = tempfile.mkstemp(dir=os.path.dirname(path))
        try:
            with os.fdopen(fd, "wb") as f:
                f.write(data)
            os.replace(tmp, path)
        except BaseException:
            if os.path.exists(tmp):
                os.unlink(tmp)
            raise

    def __getitem__(self, cid):
        try:
            with open(self._path(cid), "rb") as f:
                data = f.read()
        except FileNotFoundError:
            raise KeyError(bytes(cid)) from None              # Error::NotFound, resolver.rs:143
        if self.verify and cid_of(data) != bytes(cid):
            raise ValueError("stored object does not hash to its CID")
        return data

    def get(self, cid, default=None):
        try:
            return self[cid]
        except KeyError:
            return default

    def __iter__(self):
        for sub in sorted(os.listdir(self.root)):
            d = os.path.join(self.root, sub)
            if os.path.isdir(d):
                for name in sorted(os.listdir(d)):
                    if len(name) == 2 * CID_BYTES:
                        yield bytes.fromhex(name)

    def __len__(self):
        return sum(1 for _ in self)
