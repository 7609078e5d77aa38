"""The one entry point of nose_parameterized the reference's tests use (test/test_representation_graphs.py:1,16):
`@parameterized.expand(cases)` on a unittest.TestCase method generates one test method per case, named
<method>_<index>_<first argument>."""
import functools
import inspect
import re


class parameterized(object):
    @staticmethod
    def expand(cases):
        cases = [tuple(c) for c in cases]

        def decorator(func):
            namespace = inspect.currentframe().f_back.f_locals          # the class body being executed
            for index, args in enumerate(cases):
                suffix = re.sub(r'\W+', '_', str(args[0])) if args else ''
                name = '%s_%d_%s' % (func.__name__, index, suffix)

                def make(bound_args):
                    @functools.wraps(func)
                    def test(self):
                        return func(self, *bound_args)
                    return test

                generated = make(args)
                generated.__name__ = name
                namespace[name] = generated
            return None                                                  # the template itself is not a test

        return decorator
