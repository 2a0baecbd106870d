"""Learner half of agent.Agent (agent.py:209-273) -- SURVEY.md section 8 f-1, marked NEXT: it is
outside the rollout hot path of this round.  The API is kept so that Agent.play's
`game_step % 128 == 0 -> update_strategy()` hook resolves; calling it with enough data raises
until the fused forward+backward+SGD kernels land."""


def update_best_response(agent):
    raise NotImplementedError("update_best_response_network: learner kernels are SURVEY 8 f-1 (next round)")


def update_average_policy(agent):
    raise NotImplementedError("update_avg_response_network: learner kernels are SURVEY 8 f-1 (next round)")
