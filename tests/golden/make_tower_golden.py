"""Golden embeddings of the seeded depth-2 stand-in towers (oracle/towers.py), generated in the build container by
importing transformers (the reference's own towers -- open_clip/timm -- are not importable here).

    python tests/golden/make_tower_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import towers as T  # noqa: E402


def main():
    v = T.build_vision(depth=2, seed=42)
    t = T.build_text(depth=2, seed=43)
    imgs = T.synthetic_images(1, 2)
    ids = T.synthetic_token_ids(1, 3)
    fi, hi = T.encode_image(v, imgs, hidden_states=True)
    ft, ht = T.encode_text(t, ids, hidden_states=True)
    np.savez_compressed(os.path.join(HERE, "towers_depth2.npz"), image_features=fi, text_features=ft,
                        image_hidden_norms=np.array([np.linalg.norm(h) for h in hi], np.float32),
                        text_hidden_norms=np.array([np.linalg.norm(h) for h in ht], np.float32),
                        image_token0_block2=hi[2][0, 0], text_last_block2=ht[2][0, -1])
    print("wrote towers_depth2.npz", fi.shape, ft.shape)


if __name__ == "__main__":
    main()
