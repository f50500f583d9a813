/* oracle/katz_ref_stubs.c -- TEST INFRASTRUCTURE: abort() stubs for the mgp_* functions the online (mgp::Graph) half of
 * the reference katz.cpp imports; the static path never reaches them (see katz_ref_shim.cpp). */
#include <stdlib.h>
void mgp_edge_copy(void) { abort(); }
void mgp_edge_destroy(void) { abort(); }
void mgp_edge_get_from(void) { abort(); }
void mgp_edge_get_id(void) { abort(); }
void mgp_edge_get_to(void) { abort(); }
void mgp_edges_iterator_destroy(void) { abort(); }
void mgp_edges_iterator_get(void) { abort(); }
void mgp_edges_iterator_next(void) { abort(); }
void mgp_graph_get_vertex_by_id(void) { abort(); }
void mgp_graph_iter_vertices(void) { abort(); }
void mgp_vertex_copy(void) { abort(); }
void mgp_vertex_destroy(void) { abort(); }
void mgp_vertex_get_id(void) { abort(); }
void mgp_vertex_iter_in_edges(void) { abort(); }
void mgp_vertex_iter_out_edges(void) { abort(); }
void mgp_vertices_iterator_destroy(void) { abort(); }
void mgp_vertices_iterator_get(void) { abort(); }
void mgp_vertices_iterator_next(void) { abort(); }
