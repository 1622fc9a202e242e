from pipegoose_b200.nn.pipeline_parallel.pipeline_parallel import PipelineParallel

__all__ = ["PipelineParallel"]
