"""Clock of the routine model: which hour, day and month an iteration falls in.

Same function names and results as covid_abs/network/util.py:3-63 (one iteration = one hour).  Quirks that shape
results and are therefore kept: iteration 0 is hour 0 of day 1 of the week, days 1 and 7 are the non-working ones,
so a run starts on a day off; hour 16 of a working day matches none of bed/work/lunch/free, so nobody moves then;
hour 12 is both work time and lunch time and lunch wins in the step (graph_abs.py:213-220).  The device evaluates
the same table inside the step kernel (csrc/graph_kernels.cuh).
"""
HOURS_PER_DAY, DAYS_PER_WEEK, DAYS_PER_MONTH = 24, 7, 30
_WINDOWS = {'bed': (0, 8), 'work': (8, 16), 'free': (17, 24)}   # [start, end) hours of the day


def _hour(iteration):
    return iteration % HOURS_PER_DAY


def number_of_days(iteration):
    return iteration // HOURS_PER_DAY


def check_time(iteration, start, end):
    return start <= _hour(iteration) < end


def new_day(iteration):
    return _hour(iteration) == 0


def day_of_week(iteration):
    return number_of_days(iteration) % DAYS_PER_WEEK + 1


def work_day(iteration):
    return 1 < day_of_week(iteration) < DAYS_PER_WEEK


def day_of_month(iteration):
    return number_of_days(iteration) % DAYS_PER_MONTH + 1


def new_month(iteration):
    return new_day(iteration) and day_of_month(iteration) == 1


def bed_time(iteration):
    return check_time(iteration, *_WINDOWS['bed'])


def work_time(iteration):
    return check_time(iteration, *_WINDOWS['work'])


def lunch_time(iteration):
    return _hour(iteration) == 12


def free_time(iteration):
    return check_time(iteration, *_WINDOWS['free'])
