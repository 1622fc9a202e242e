from pipegoose_b200.core.bucket.bucket import Bucket
from pipegoose_b200.core.bucket.dist import BucketDistributor
from pipegoose_b200.core.bucket.manager import BucketManager

__all__ = ["Bucket", "BucketDistributor", "BucketManager"]
