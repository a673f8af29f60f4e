"""Edward2 interceptor stack (TFP 0.7 edward2/interceptor.py semantics): the most
recently pushed interceptor runs first and is popped while it runs, so nested
`interceptable` calls reach the next one out."""
import contextlib
import functools
import threading
from collections import OrderedDict


class _Stack(threading.local):
    def __init__(self):
        super(_Stack, self).__init__()
        self.stack = [lambda f, *a, **k: f(*a, **k)]


_interceptor_stack = _Stack()


@contextlib.contextmanager
def interception(interceptor):
    try:
        _interceptor_stack.stack.append(interceptor)
        yield
    finally:
        _interceptor_stack.stack.pop()


@contextlib.contextmanager
def get_next_interceptor():
    try:
        interceptor = _interceptor_stack.stack.pop()
        yield interceptor
    finally:
        _interceptor_stack.stack.append(interceptor)


def interceptable(func):
    @functools.wraps(func)
    def func_wrapped(*args, **kwargs):
        with get_next_interceptor() as interceptor:
            return interceptor(func, *args, **kwargs)
    return func_wrapped


@contextlib.contextmanager
def tape():
    tape_data = OrderedDict()

    def record(f, *args, **kwargs):
        name = kwargs.get("name")
        output = interceptable(f)(*args, **kwargs)
        if name:
            tape_data[name] = output
        return output

    with interception(record):
        yield tape_data
