// std::sort as libstdc++ implements it (bits/stl_algo.h: introsort with a median-of-3 pivot,
// threshold 16, heapsort fallback, then one final insertion pass), restated over an int32 array
// with a caller-supplied strict-weak `less(a, b)` on the stored values.
//
// std::sort is not stable and the reference sorts lists that contain ties (starpilot.cpp:356 sorts
// its spawners by spawn_time only), so the ORDER OF TIES is part of the behaviour to reproduce:
// the sequence of comparisons and swaps below is the library's, step for step. The oracle is
// built with the same libstdc++; tests/test_oracle.py checks this restatement against it.
#pragma once
#include "pg_common.cuh"

namespace pg {

template <class Less>
struct StdSort {
    int32_t *a;
    Less less;
    PG_HD StdSort(int32_t *a_, Less l) : a(a_), less(l) {}

    PG_HD void swap_at(int i, int j) {
        int32_t t = a[i];
        a[i] = a[j];
        a[j] = t;
    }
    // __move_median_to_first
    PG_HD void move_median_to_first(int result, int x, int y, int z) {
        if (less(a[x], a[y])) {
            if (less(a[y], a[z]))
                swap_at(result, y);
            else if (less(a[x], a[z]))
                swap_at(result, z);
            else
                swap_at(result, x);
        } else if (less(a[x], a[z])) {
            swap_at(result, x);
        } else if (less(a[y], a[z])) {
            swap_at(result, z);
        } else {
            swap_at(result, y);
        }
    }
    // __unguarded_partition
    PG_HD int unguarded_partition(int first, int last, int pivot) {
        while (true) {
            while (less(a[first], a[pivot])) ++first;
            --last;
            while (less(a[pivot], a[last])) --last;
            if (!(first < last))
                return first;
            swap_at(first, last);
            ++first;
        }
    }
    // __adjust_heap followed by __push_heap, on the heap that starts at `first`
    PG_HD void adjust_heap(int first, int hole, int len, int32_t value) {
        const int top = hole;
        int child = hole;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (less(a[first + child], a[first + (child - 1)]))
                child--;
            a[first + hole] = a[first + child];
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            a[first + hole] = a[first + (child - 1)];
            hole = child - 1;
        }
        int parent = (hole - 1) / 2;
        while (hole > top && less(a[first + parent], value)) {
            a[first + hole] = a[first + parent];
            hole = parent;
            parent = (hole - 1) / 2;
        }
        a[first + hole] = value;
    }
    // __partial_sort(first, last, last): __heap_select (= __make_heap here) then __sort_heap
    PG_HD void heap_sort(int first, int last) {
        const int len = last - first;
        if (len >= 2) {
            int parent = (len - 2) / 2;
            while (true) {
                int32_t v = a[first + parent];
                adjust_heap(first, parent, len, v);
                if (parent == 0)
                    break;
                parent--;
            }
        }
        while (last - first > 1) {
            --last;
            int32_t v = a[last];  // __pop_heap(first, last, last)
            a[last] = a[first];
            adjust_heap(first, 0, last - first, v);
        }
    }
    // __unguarded_linear_insert
    PG_HD void unguarded_linear_insert(int last) {
        int32_t val = a[last];
        int next = last - 1;
        while (less(val, a[next])) {
            a[last] = a[next];
            last = next;
            --next;
        }
        a[last] = val;
    }
    // __insertion_sort
    PG_HD void insertion_sort(int first, int last) {
        if (first == last)
            return;
        for (int i = first + 1; i != last; ++i) {
            if (less(a[i], a[first])) {
                int32_t val = a[i];
                for (int k = i; k > first; k--) a[k] = a[k - 1];
                a[first] = val;
            } else {
                unguarded_linear_insert(i);
            }
        }
    }
    // std::sort(first, last): the recursion of __introsort_loop on the right-hand part is unrolled
    // onto a small explicit stack (depth <= 2*lg(n) by construction)
    PG_HD void sort(int n) {
        if (n <= 0)
            return;
        int lg = 0;
        while ((n >> (lg + 1)) != 0) lg++;
        int stack_first[64], stack_last[64], stack_depth[64];
        int sp = 0;
        stack_first[0] = 0;
        stack_last[0] = n;
        stack_depth[0] = lg * 2;
        sp = 1;
        while (sp > 0) {
            sp--;
            int first = stack_first[sp];
            int last = stack_last[sp];
            int depth = stack_depth[sp];
            // one activation of __introsort_loop(first, last, depth): the recursive calls on
            // [cut, last) run BEFORE the loop continues on [first, cut); they touch disjoint
            // ranges, so deferring the left part and doing the right part first is equivalent.
            while (last - first > 16) {
                if (depth == 0) {
                    heap_sort(first, last);
                    break;
                }
                --depth;
                const int mid = first + (last - first) / 2;
                move_median_to_first(first, first + 1, mid, last - 1);
                const int cut = unguarded_partition(first + 1, last, first);
                // push right part [cut, last) with the decremented depth
                stack_first[sp] = cut;
                stack_last[sp] = last;
                stack_depth[sp] = depth;
                sp++;
                last = cut;
            }
        }
        // __final_insertion_sort
        if (n > 16) {
            insertion_sort(0, 16);
            for (int i = 16; i != n; ++i) unguarded_linear_insert(i);
        } else {
            insertion_sort(0, n);
        }
    }
};

template <class Less>
PG_HD void pg_std_sort(int32_t *a, int n, Less less) {
    StdSort<Less> s(a, less);
    s.sort(n);
}

}  // namespace pg
