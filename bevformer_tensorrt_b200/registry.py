"""Name -> function registry with the reference's interface (det2trt/models/utils/register.py:9-69):
``register_module(name=None, force=False, module=None)`` usable directly or as a decorator, ``get(key)``,
``module_dict``. Modules bind their op once at construction time through it (spatial_cross_attention.py:692)."""


class FuncRegistry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    name = property(lambda self: self._name)
    module_dict = property(lambda self: self._module_dict)

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return key in self._module_dict

    def __repr__(self):
        return f"{type(self).__name__}(name={self._name}, items={sorted(self._module_dict)})"

    def get(self, key):
        return self._module_dict.get(key)

    def _add(self, module, name, force):
        names = [module.__name__] if name is None else ([name] if isinstance(name, str) else list(name))
        for n in names:
            if not force and n in self._module_dict:
                raise KeyError(f"{n} is already registered in {self._name}")
            self._module_dict[n] = module

    def register_module(self, name=None, force=False, module=None):
        if not isinstance(force, bool):
            raise TypeError(f"force must be a boolean, but got {type(force)}")
        if not (name is None or isinstance(name, str)):
            raise TypeError(f"name must be None or a str, but got {type(name)}")
        if module is not None:
            self._add(module, name, force)
            return module

        def decorator(fn):
            self._add(fn, name, force)
            return fn

        return decorator


TRT_FUNCTIONS = FuncRegistry("tensorrt functions")
