"""Namespace filled by quimb/tensor/__init__.py (stand-in)."""
