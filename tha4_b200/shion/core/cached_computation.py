"""Memoised DAG evaluation used by every poser mode -- mirror of src/tha4/shion/core/cached_computation.py:9-59."""
from abc import ABC, abstractmethod
from typing import Any, Callable, Dict, Optional

from torch.nn import Module


class ComputationState:   # cached_computation.py:9-20
    def __init__(self, modules: Dict[str, Module], accumulated_modules: Dict[str, Module], batch: Any,
                 outputs: Optional[Dict[str, Any]] = None):
        self.outputs = {} if outputs is None else outputs
        self.batch = batch
        self.accumulated_modules = accumulated_modules
        self.modules = modules


CachedComputationFunc = Callable[[ComputationState], Any]


class CachedComputationProtocol(ABC):   # cached_computation.py:41-59
    def get_output(self, key: str, state: ComputationState) -> Any:
        if key not in state.outputs:
            state.outputs[key] = self.compute_output(key, state)
        return state.outputs[key]

    @abstractmethod
    def compute_output(self, key: str, state: ComputationState) -> Any:
        pass

    def get_output_func(self, key: str) -> CachedComputationFunc:
        def func(state: ComputationState):
            return self.get_output(key, state)

        return func
