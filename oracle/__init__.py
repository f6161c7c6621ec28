"""CPU oracle for the UnCLTMO hot path.  TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (CPU, fp32 or fp64) restatement of the reference algorithm for every row of
SURVEY.md §8(a).  Each function cites the reference file:line it follows.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may
import this package; the product package `uncltmo_b200` never does.

Pinning: the reference ships no golden vectors, so the oracle is pinned against
outputs of the reference itself, imported in place on CPU by `tests/golden/make_golden.py`
(fixtures committed under `tests/golden/*.npz`, checked by `tests/test_oracle_golden.py`).
"""
from .generator import (unet_forward, unet_video_forward, gcn_block, relative_pos_table,  # noqa: F401
                        knn_indices, bf16_operands)
from .discriminator import simple_discriminator_forward, contrast_map, gauss_window  # noqa: F401
from .losses import (struct_loss, contrastive_d_loss, nce, l1_mean_terms, tv_loss)  # noqa: F401
from .frame_path import (log_lambda_normalise, to_gray, resize_im, tile_and_blend, tile_grid,  # noqa: F401
                         back_to_color, postprocess_frame, tonemap_frame)
from .train_step import tmqi_naturalness, train_step_losses  # noqa: F401
