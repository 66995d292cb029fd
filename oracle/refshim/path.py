"""the two `path.py` methods the reference's GANcheckpoints.save_weights uses."""
import os


class Path(str):
    def exists(self): return os.path.exists(self)
    def stripext(self): return Path(os.path.splitext(self)[0])
    def rename(self, new): os.rename(self, new); return Path(new)
