"""matplotlib stand-in (import-only; debug plotting is out of scope)."""
