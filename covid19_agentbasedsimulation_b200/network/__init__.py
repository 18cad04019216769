"""Mirror of `covid_abs.network`: the hourly routine model (GraphSimulation) on the B200."""
