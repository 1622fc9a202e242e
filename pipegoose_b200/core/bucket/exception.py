class BucketFullError(Exception):
    """The tensor does not fit into the remaining space of the bucket."""


class BucketClosedError(Exception):
    """The bucket was closed and accepts no more tensors."""
