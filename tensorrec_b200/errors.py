"""Exception types of the TensorRec API.  Class names, constructor keywords and message texts are part of the
reference's error contract (tensorrec/errors.py:4-41: callers catch these classes and tests match the texts), so they
are reproduced exactly; the classes are generated from the table below."""


class TensorRecException(Exception):
    """Base class: `msg` is a str.format template filled from the constructor's keyword arguments."""
    msg = None

    def __init__(self, **kwargs):
        Exception.__init__(self)
        self.kwargs = kwargs

    def __str__(self):
        return self.msg.format(**self.kwargs)

    __repr__ = __str__

    @property
    def message(self):
        return str(self)


_ERROR_TABLE = (
    # (class name, message template, note)
    ('ModelNotBiasedException', 'Cannot predict {actor} bias for unbiased model',
     'predict_user_bias / predict_item_bias on a model built with biased=False'),
    ('ModelNotFitException',
     '{method}() has been called before model fitting. Call fit() or fit_partial() before calling {method}().',
     'any predict* method before the first fit'),
    ('ModelWithoutAttentionException',
     "This TensorRec model does not use attention. Try re-building TensorRec with a valid 'attention_graph' arg.",
     'predict_user_attention_representation on a model without an attention graph'),
    ('BatchNonSparseInputException',
     'In order to support user batching at fit time, interactions and user_features must both be scipy.sparse '
     'matrices.', 'user_batch_size with Dataset / TFRecord inputs'),
    ('TfVersionException',
     'You need to have at least TensorFlow version 1.7 installed in order to use TensorRec properly. You have '
     'currently installed TensorFlow: {tf_version}',
     'kept for API compatibility; this build has no TensorFlow to version-check'),
)

for _name, _template, _note in _ERROR_TABLE:
    globals()[_name] = type(_name, (TensorRecException,), {'msg': _template, '__doc__': _note,
                                                            '__module__': __name__})
del _name, _template, _note
