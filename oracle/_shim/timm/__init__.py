"""Test-only shim: `timm.utils.ModelEmaV2` is only a deepcopy holder at inference (model.py:3657)."""
