"""No-op SummaryWriter with the calls the reference's train scripts make (add_scalar / add_image / close)."""


class SummaryWriter:
    def __init__(self, *a, **k):
        self.scalars = 0

    def add_scalar(self, tag, value, step=None, *a, **k):
        float(value)            # the scripts pass 0-dim CUDA tensors: force the same host sync the real writer does
        self.scalars += 1

    def add_image(self, *a, **k):
        pass

    def close(self):
        pass
