"""gym.error stand-in (test infrastructure only)."""


class Error(Exception):
    pass
