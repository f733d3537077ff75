"""tf.keras.regularizers."""


class L2(object):
    def __init__(self, l=0.01):      # noqa: E741
        self.l2 = float(l)


def l2(l=0.01):                      # noqa: E741
    return L2(l)
