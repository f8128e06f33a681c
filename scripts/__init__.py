"""Reference-compatible import path."""
