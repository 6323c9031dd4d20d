import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lenslesspicam_b200 as L
g = os.path.join(ROOT, "tests", "golden")
psf, data = L.load_data(os.path.join(g, "raw_psf.png"), os.path.join(g, "raw_data.png"), downsample=1.6, normalize=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "loader_dump.npz"), psf=psf.cpu().numpy(), data=data.cpu().numpy())
