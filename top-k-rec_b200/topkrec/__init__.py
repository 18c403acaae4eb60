"""ctypes binding of libtopkrec.so (the sm_100a engine; ``include/topkrec.h``).

PyTorch is only the carrier of device buffers and the current stream: every
entry point receives raw ``data_ptr()`` values.  There is no CPU or eager
fallback -- importing this module without the built library, or calling it
without a CUDA device, raises.
"""
from ._lib import (TkrError, BprCfg, VbprCfg, Sampler, vbpr_workspace, vbpr_set_hot_items, vbpr_project, vbpr_step, vbpr_grad, vbpr_apply, vbpr_grad_views, lib, version, launch_count, reset_launch_count,  # noqa: F401
                   bpr_workspace, bpr_workspace_layout, bpr_set_hot_items, popular_items, MAX_HOT, bpr_item_grad_view, bpr_grad, bpr_apply, bpr_dp_layout, bpr_dp_step, bpr_dp_status, bpr_step, bpr_hogwild, bpr_step_host, bpr_sample, score_topk, score_topk_segment, score_topk_host, score_topk_batches, BatchScorer, topk_merge, eval_hits, dat_read, dat_write, ratings_parse)
from .als import AlsSide, build_plan as als_build_plan, als_gram, als_solve_rows  # noqa: F401,E402
