"""Compare two results.pkl files: times, solution lengths, nodes generated, nodes/sec, % equal length
(the reference's scripts/compare_solutions.py; Nodes/Sec here is the metric BASELINE.json quotes)."""
import os
import pickle
import sys
from argparse import ArgumentParser

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def print_stats(data, hist=False):
    data = np.asarray(data, dtype=np.float64)
    print("Min/Max/Median/Mean(Std) %f/%f/%f/%f(%f)" % (data.min(), data.max(), float(np.median(data)), float(data.mean()),
                                                        float(data.std())))
    if hist:
        counts, edges = np.histogram(data)
        for c, e in zip(counts, edges):
            print("%s %s" % (c, e))


def print_results(results):
    times = np.array(results["times"])
    lens = np.array([len(x) for x in results["solutions"]])
    nodes = np.array(results["num_nodes_generated"])
    for title, arr in (("-Times-", times), ("-Lengths-", lens), ("-Nodes Generated-", nodes), ("-Nodes/Sec-", nodes / times)):
        print(title)
        print_stats(arr)


def main():
    parser = ArgumentParser()
    parser.add_argument("--soln1", type=str, required=True)
    parser.add_argument("--soln2", type=str, required=True)
    args = parser.parse_args()
    r1 = pickle.load(open(args.soln1, "rb"))
    r2 = pickle.load(open(args.soln2, "rb"))
    lens1 = np.array([len(x) for x in r1["solutions"]])
    lens2 = np.array([len(x) for x in r2["solutions"]])
    print("%i states" % len(r1["states"]))
    print("\n--SOLUTION 1---")
    print_results(r1)
    print("\n--SOLUTION 2---")
    print_results(r2)
    print("\n\n------Solution 2 - Solution 1 Lengths-----")
    print_stats(lens2 - lens1)
    print("%.2f%% soln2 equal to soln1" % (100 * np.mean(lens2 == lens1)))


if __name__ == "__main__":
    main()
