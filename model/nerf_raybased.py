"""Drop-in for the reference's `model/nerf_raybased.py`: same import path, same names.

    from model.nerf_raybased import NeRF, NeRF_v3_2, PositionalEmbedder, PointSampler   # main.py:13
    from model.nerf_raybased import NeRF                                                # utils/create_data.py:10

Everything is implemented in r2l_b200/ (B200-native kernels behind the C ABI); this file only re-exports.
Pickled checkpoints that name `model.nerf_raybased.NeRF_v3_2` (main.py:1534-1536) resolve to the classes here.
"""
from r2l_b200.nerf_raybased import *  # noqa: F401,F403
from r2l_b200.nerf_raybased import (Embedder, EncodedPoints, NeRF, NeRF_v3_2, PointSampler, PositionalEmbedder, ResMLP, batchify, device,  # noqa: F401
                                    get_activation, get_embedder, img2mse, mse2psnr, raw2outputs, run_network, to8b,
                                    to_array, to_list, to_tensor)
