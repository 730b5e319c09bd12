"""Exception classes (mirror of librosa/util/exceptions.py:1-15)."""


class LibrosaError(Exception):
    """Root of the exception hierarchy used on this path."""


class ParameterError(LibrosaError):
    """Raised for malformed or out-of-range arguments, under the same conditions as librosa."""
