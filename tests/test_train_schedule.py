"""The outer training schedule (train_Enero_3top.py:28-34, train_Enero_3top_script.py:428-444) against the reference's
own log: Logs/expSP_3top_15_B_NEWLogs.txt prints one `lr,` line per PPO episode -- 60 x 0.0002, 60 x 0.000192,
60 x 0.0001769472, 60 x 0.0001565515579392, 60 x 0.0001329665271983002, 60 x 0.00010841727597218178, then 0.0001
(the floor) -- i.e. the decay compounds on the value pickled by the previous launch."""
import itertools

from enero_b200.train import launch_schedule

LOGGED = [0.0002] * 60 + [0.000192] * 60 + [0.0001769472] * 60 + [0.0001565515579392] * 60 + \
         [0.0001329665271983002] * 60 + [0.00010841727597218178] * 60 + [0.0001] * 240


def test_learning_rate_sequence_matches_the_reference_log():
    got = [lr for _, lr, _, _ in launch_schedule(0, len(LOGGED), 0.0002)]
    assert got == LOGGED


def test_schedule_resumes_from_the_pickled_rate_and_entropy_drops_at_60():
    # a launch that starts at iteration 120 with the pickled 0.000192 continues the same sequence
    got = [lr for _, lr, _, _ in launch_schedule(120, 60, 0.000192)]
    assert got == LOGGED[120:180]
    betas = [b for _, _, b, _ in launch_schedule(0, 120, 0.0002)]
    assert betas[:60] == [0.01] * 60 and betas[60:] == [0.001] * 60
    bounds = [i for i, _, _, b in launch_schedule(0, 100, 0.0002) if b]
    assert bounds == [0, 20, 40, 60, 80]
