"""Minimal stand-in for gym 0.17 (absent from this image), only what the reference imports.
Test infrastructure for oracle/ref_harness.py; never imported by the product."""
from . import error, spaces, utils, envs  # noqa: F401


class Env:
    metadata = {}
