"""B200-native batched MCTS engine for the search hot path of timoklein/alphazero-gym.

Public surface (mirrors the reference's):
    alphazero_gym_b200.search.mcts.MCTSDiscrete / MCTSContinuous   drop-in for alphazero.search.mcts
    alphazero_gym_b200.agent.agents.DiscreteAgent / ContinuousAgent  act() over the engine
    alphazero_gym_b200.engine.SearchEngine                          batched trees on one GPU
The compute path is hand-written CUDA (csrc/, sm_100a) behind the C ABI in include/azg.h.
"""
__version__ = "0.1.0"
