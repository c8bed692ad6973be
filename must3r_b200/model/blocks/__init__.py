"""Import paths of must3r/model/blocks kept for callers that reach into them (only the attention toggles are used outside
the model package: must3r/model/blocks/attention.py:18-27)."""
