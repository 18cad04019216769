"""Helpers shared by the GraphSimulation tests: oracle `World` dict <-> product columns, scenario kwargs."""
import numpy as np

PERSON = ("px", "py", "status", "severity", "infected_time", "age", "stratum", "active", "isolated", "house", "employer",
          "hire_seq", "pwealth")
HOUSE = ("hx", "hy", "hstratum", "hsize", "hwealth", "hfixed")
BUSINESS = ("bx", "by", "bstratum", "bstocks", "bnum_employees", "bprice", "bwealth", "bfixed")


def product_world(ow):
    """oracle World (oracle/graph_oracle.py) -> columns of cabs_graph_world."""
    w = {k: np.asarray(ow[k]) for k in PERSON + HOUSE + BUSINESS}
    g, h = ow["gov"], ow["health"]
    w.update(gov_x=int(g["x"]), gov_y=int(g["y"]), gov_wealth=float(g["wealth"]), gov_price=float(g["price"]),
             gov_fixed=float(g["fixed"]), health_x=int(h["x"]), health_y=int(h["y"]), health_wealth=float(h["wealth"]),
             health_fixed=float(h["fixed"]), health_expenses=float(h.get("expenses", 0.0)), iteration=-1)
    return w


def oracle_world(pw, params):
    """columns of cabs_graph_world (+ oracle params dict) -> oracle World."""
    ow = {k: np.asarray(pw[k]) for k in PERSON + HOUSE + BUSINESS}
    ow["gov"] = dict(x=pw["gov_x"], y=pw["gov_y"], wealth=pw["gov_wealth"], price=pw["gov_price"], fixed=pw["gov_fixed"])
    ow["health"] = dict(x=pw["health_x"], y=pw["health_y"], wealth=pw["health_wealth"], fixed=pw["health_fixed"],
                        expenses=pw["health_expenses"])
    ow["params"] = dict(params)
    return ow


def sim_kwargs(params):
    """GraphSimulation kwargs from the params block of an oracle World."""
    from covid19_agentbasedsimulation_b200.agents import Status
    amp = params["amplitudes"]
    return dict(population_size=params["population_size"], total_wealth=params["total_wealth"],
                contagion_distance=params["contagion_distance"], contagion_rate=params["contagion_rate"],
                critical_limit=params["critical_limit"],
                amplitudes={Status.Susceptible: amp[0], Status.Infected: amp[1], Status.Recovered_Immune: amp[2]},
                minimum_income=params["minimum_income"], minimum_expense=params["minimum_expense"],
                business_distance=params["business_distance"], incubation_time=params["incubation_time"],
                contagion_time=params["contagion_time"], recovering_time=params["recovering_time"],
                money_bits=params["money_bits"])
