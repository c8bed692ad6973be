from .inference import (postprocess, stack_views, encoder_multi_ar, inference_multi_ar_batch, inference_multi_ar,  # noqa: F401
                        inference_video_multi_ar, inference_encoder, inference, get_Nmem, unstack_pointmaps,
                        concat_preds, groupby_consecutive)
from .memory_io import plain_memory, save_memory, load_memory  # noqa: F401,E402
