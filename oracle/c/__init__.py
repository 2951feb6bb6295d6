"""C restatement of the hot path (TEST INFRASTRUCTURE; see oracle/__init__.py)."""
