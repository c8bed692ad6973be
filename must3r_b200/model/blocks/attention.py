"""must3r/model/blocks/attention.py:5-27: module-level switches callers flip at start-up; see must3r_b200.compat.attention."""
from ...compat.attention import (has_xformers, has_scaled_dot_product_attention, toggle_memory_efficient_attention,  # noqa: F401
                                 is_memory_efficient_attention_enabled, attention)
