"""Stand-in for the `pymoo` package (not installable offline), covering
exactly what the reference's scheduling policy uses
(``sched/adaptdl_sched/policy/pollux.py``): the ``Problem`` / ``Crossover`` /
``Mutation`` / ``Repair`` base classes, ``NSGA2`` with a sampled start
population, ``minimize(problem, algorithm, ("n_gen", N))`` and
``NonDominatedSorting``. Written for this repository's baseline arm; it is
an ordinary NSGA-II (binary tournaments on rank and crowding distance, 90 %
crossover probability, duplicate elimination, elitist rank-and-crowding
survival, result = the non-dominated set of the final population), NOT a copy
of pymoo. With it the UNMODIFIED reference policy can be run next to this
framework's (``tools/sched_sim.py --policy reference``,
``tools/policy_bench.py --search reference``).
"""
