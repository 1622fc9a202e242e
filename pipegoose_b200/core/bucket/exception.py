"""Errors of the packing buffers (parity: reference core/bucket/exception.py).  Both derive from :class:`BucketError`
so that callers which only want to "start a new bucket and retry" can catch one type."""


class BucketError(RuntimeError):
    pass


class BucketFullError(BucketError):
    """The tensor does not fit into the space that is left; flush the bucket and open a new one."""


class BucketClosedError(BucketError):
    """``close()`` was called (the bucket is being reduced or was handed over); nothing can be added."""
