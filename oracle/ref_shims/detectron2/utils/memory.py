def retry_if_cuda_oom(func):
    return func
