"""tf.keras.callbacks: names only (trainer.py imports them; nothing here trains)."""


class _Callback(object):
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs


class ModelCheckpoint(_Callback):
    pass


class TensorBoard(_Callback):
    pass


class LearningRateScheduler(_Callback):
    pass
