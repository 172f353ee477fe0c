"""The reference's broker tests, restated 1:1 so that the same scenario runs against the oracle
(CPU) and against the CUDA engine through the C ABI (GPU).  Every function cites the reference
test it replays; assertions follow the reference's `assert_received!` calls line by line.
"""
from harness import (DA, GLOBAL, Broadcast, Direct, TestBroker, TestDefinition, TestUser, at_index)


def _bcast_definition():
    # cdn-broker/src/tests/broadcast.rs:29-49 (same in :104-124)
    return TestDefinition(
        connected_users=[TestUser(0, [GLOBAL, DA]), TestUser(1, [DA]), TestUser(2, [GLOBAL])],
        connected_brokers=[
            TestBroker([TestUser(3, [DA])]),
            TestBroker([TestUser(4, [GLOBAL, DA])]),
            TestBroker([TestUser(5, [])]),
        ],
    )


def test_broadcast_user(backend):
    """cdn-broker/src/tests/broadcast.rs:26-95"""
    run = _bcast_definition().into_run(backend)
    message = Broadcast([GLOBAL], b"test broadcast global")
    assert run.send_as_user(0, message) == 0
    run.assert_received(run.connected_users[0], message)
    run.assert_received(run.connected_users[2], message)
    run.assert_received(run.connected_brokers[1], message)
    run.assert_nothing_more()

    message = Broadcast([DA], b"test broadcast DA")
    assert run.send_as_user(2, message) == 0
    run.assert_received(run.connected_users[0], message)
    run.assert_received(run.connected_users[1], message)
    run.assert_received(run.connected_brokers[0], message)
    run.assert_received(run.connected_brokers[1], message)
    run.assert_nothing_more()


def test_broadcast_broker(backend):
    """cdn-broker/src/tests/broadcast.rs:101-167"""
    run = _bcast_definition().into_run(backend)
    message = Broadcast([GLOBAL], b"test broadcast global")
    assert run.send_as_broker(2, message) == 0
    run.assert_received(run.connected_users[0], message)
    run.assert_received(run.connected_users[2], message)
    run.assert_nothing_more()

    message = Broadcast([DA], b"test broadcast DA.")
    assert run.send_as_broker(1, message) == 0
    run.assert_received(run.connected_users[0], message)
    run.assert_received(run.connected_users[1], message)
    run.assert_nothing_more()


def test_direct_user_to_user(backend):
    """cdn-broker/src/tests/direct.rs:27-82"""
    run = TestDefinition(
        connected_users=[TestUser(0, [GLOBAL]), TestUser(1, [DA])],
        connected_brokers=[TestBroker([TestUser(2, [DA])]), TestBroker([TestUser(3, [])]),
                           TestBroker([TestUser(4, [])])],
    ).into_run(backend)
    message = Direct(at_index(0), b"test direct 0")
    assert run.send_as_user(0, message) == 0
    run.assert_received(run.connected_users[0], message)
    run.assert_nothing_more()

    message = Direct(at_index(1), b"test direct 1")
    assert run.send_as_user(1, message) == 0
    run.assert_received(run.connected_users[1], message)
    run.assert_nothing_more()


def _direct_definition():
    # cdn-broker/src/tests/direct.rs:91-107 (same in :136-152)
    return TestDefinition(
        connected_users=[TestUser(0, [GLOBAL]), TestUser(1, [GLOBAL, DA])],
        connected_brokers=[TestBroker([TestUser(2, [])]), TestBroker([TestUser(3, [DA])]),
                           TestBroker([TestUser(4, [])])],
    )


def test_direct_user_to_broker(backend):
    """cdn-broker/src/tests/direct.rs:88-127"""
    run = _direct_definition().into_run(backend)
    message = Direct(at_index(2), b"test direct 2")
    assert run.send_as_user(0, message) == 0
    run.assert_received(run.connected_brokers[0], message)
    run.assert_nothing_more()


def test_direct_broker_to_user(backend):
    """cdn-broker/src/tests/direct.rs:133-173 — a broker-origin direct for a user owned by
    another broker is never bounced"""
    run = _direct_definition().into_run(backend)
    message = Direct(at_index(2), b"test direct 2")
    assert run.send_as_broker(1, message) == 0
    run.assert_nothing_more()


def test_subscribe(backend):
    """tests/src/tests/subscribe.rs:19-121 — one client on one broker: receive own Global
    broadcast, not DA; subscribe DA → receive; unsubscribe DA → not any more.  (The client/marshal
    hop is outside the path; the client's frames enter user_receive_loop as here.)"""
    from oracle import oracle as orc

    run = TestDefinition(connected_users=[TestUser(0, [GLOBAL])]).into_run(backend)
    me = run.connected_users[0]
    m = Broadcast([GLOBAL], b"hello global")
    assert run.send_as_user(0, m) == 0
    run.assert_received(me, m)
    m = Broadcast([DA], b"hello DA")
    assert run.send_as_user(0, m) == 0
    run.assert_nothing_more()
    assert run.send_as_user(0, orc.serialize(orc.KIND_SUBSCRIBE, bytes([DA]))) == 0
    assert run.send_as_user(0, m) == 0
    run.assert_received(me, m)
    assert run.send_as_user(0, orc.serialize(orc.KIND_UNSUBSCRIBE, bytes([DA]))) == 0
    assert run.send_as_user(0, m) == 0
    run.assert_nothing_more()


def test_invalid_subscribe(backend):
    """tests/src/tests/subscribe.rs:123-197 — subscribing / unsubscribing / broadcasting with only
    invalid topics makes Topic::prune fail (def.rs:36-49), the receive loop returns Err and the
    broker drops the user (user/handler.rs:61-69)."""
    from oracle import oracle as orc

    for kind in (orc.KIND_SUBSCRIBE, orc.KIND_UNSUBSCRIBE):
        run = TestDefinition(connected_users=[TestUser(0, [GLOBAL])]).into_run(backend)
        rc = run.send_as_user(0, orc.serialize(kind, bytes([99])))
        assert rc < 0
        # the caller (handle_user_connection) removes the user on Err
        backend.remove_user(at_index(0))
        assert backend.num_users() == 0
        # nothing is delivered to the removed user afterwards
        other = backend.add_user(at_index(7), [GLOBAL])
        m = Broadcast([GLOBAL], b"after kick")
        assert backend.user_receive(at_index(7), m) == 0
        run.connected_users.append(other)
        run.assert_received(other, m)
        run.assert_nothing_more()
        backend.remove_user(at_index(7))


def test_double_connect_same_broker(backend):
    """tests/src/tests/double_connect.rs:16-58 — a second connection with the same key kicks the
    first (Connections::add_user → remove_user, connections/mod.rs:289-290): only the new
    connection receives."""
    run = TestDefinition(connected_users=[TestUser(0, [GLOBAL])]).into_run(backend)
    old = run.connected_users[0]
    new = backend.add_user(at_index(0), [GLOBAL])
    # (the engine may hand the kicked connection's dense id to the new connection: ids are only
    # meaningful between add and remove, like a slab index)
    if new != old:
        run.connected_users.append(new)
    m = Direct(at_index(0), b"hello direct")
    assert run.send_as_user(0, m) == 0
    run.assert_received(new, m)
    run.assert_nothing_more()
    m = Broadcast([GLOBAL], b"hello again")
    assert run.send_as_user(0, m) == 0
    run.assert_received(new, m)
    run.assert_nothing_more()


def test_user_moved_to_other_broker(backend):
    """tests/src/tests/double_connect.rs:60-141 seen from the first broker: a UserSync saying the
    user now lives on another broker (higher version) removes the local user
    (Connections::apply_user_sync, connections/mod.rs:154-162) and directs are forwarded there."""
    run = TestDefinition(
        connected_users=[TestUser(0, [GLOBAL]), TestUser(1, [GLOBAL])],
        connected_brokers=[TestBroker([])],
    ).into_run(backend)
    backend.apply_user_sync("0/0", [(at_index(0), 2, "0/0")])
    assert backend.num_users() == 1
    m = Direct(at_index(0), b"to the moved user")
    assert run.send_as_user(1, m) == 0
    run.assert_received(run.connected_brokers[0], m)
    run.assert_nothing_more()
    m = Broadcast([GLOBAL], b"global after move")
    assert run.send_as_user(1, m) == 0
    run.assert_received(run.connected_users[1], m)
    run.assert_nothing_more()


def test_multi_topic_dedup(backend):
    """R3: a user subscribed to several of a message's topics gets ONE copy
    (HashSet in get_interested_by_topic, connections/mod.rs:100-108); duplicate topics on the wire
    are only deduplicated when consecutive (Topic::prune dedup(), def.rs:38)."""
    run = TestDefinition(
        connected_users=[TestUser(0, [GLOBAL, DA]), TestUser(1, [DA]), TestUser(2, [])],
        connected_brokers=[TestBroker([TestUser(3, [GLOBAL, DA])])],
    ).into_run(backend)
    m = Broadcast([GLOBAL, DA, GLOBAL], b"both topics")
    assert run.send_as_user(2, m) == 0
    run.assert_received(run.connected_users[0], m)
    run.assert_received(run.connected_users[1], m)
    run.assert_received(run.connected_brokers[0], m)
    run.assert_nothing_more()


def test_fifo_order(backend):
    """R9: per connection, frames arrive in the order the broker handled them."""
    run = TestDefinition(connected_users=[TestUser(0, [GLOBAL]), TestUser(1, [GLOBAL, DA])]).into_run(backend)
    msgs = []
    for i in range(12):
        if i % 3 == 0:
            m = Direct(at_index(1), b"d%d" % i)
        elif i % 3 == 1:
            m = Broadcast([GLOBAL], b"g%d" % i * (i + 1))
        else:
            m = Broadcast([DA], b"a%d" % i)
        msgs.append(m)
        assert run.send_as_user(0, m) == 0
    for i, m in enumerate(msgs):
        run.assert_received(run.connected_users[1], m)
        if i % 3 == 1:
            run.assert_received(run.connected_users[0], m)
    run.assert_nothing_more()


ALL = [
    test_broadcast_user, test_broadcast_broker, test_direct_user_to_user, test_direct_user_to_broker,
    test_direct_broker_to_user, test_subscribe, test_invalid_subscribe, test_double_connect_same_broker,
    test_user_moved_to_other_broker, test_multi_topic_dedup, test_fifo_order,
]
